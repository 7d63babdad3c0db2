"""Scene ingestion: the product's file readers (include/ptb_sceneio.h, pathtracer_b200/csrc/scene_io.cpp)
against what the REFERENCE's own readers leave in memory for the fixtures under tests/golden/assets/ — committed in
tests/golden/sceneio.npz by tests/golden/make_golden.py, and live against oracle/_ref when it is present — and, end to end, the
oracle rendering a scene loaded by the product's reader against the reference rendering the same file loaded by its own
`Raytracer::load_scene`.  Integer / byte / index data and parsed floats must be identical; the images bit-identical (CPU) or
within the GPU tolerances of tests/parity_cases.py."""
import ctypes as C
import json
import os

import numpy as np
import pytest
import sceneio_cases as sio
from parity_cases import check_images

import pathtracer_b200
from pathtracer_b200 import _abi, api

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sceneio.npz")


@pytest.fixture(scope="module")
def io():
    return pathtracer_b200.sceneio()        # host code of libptb200.so: loads and runs without a GPU


@pytest.fixture(scope="module")
def gold():
    g = np.load(GOLD)
    return g, json.loads(bytes(g["meta_json"]).decode())


def _norm(d):
    return json.loads(json.dumps(d, sort_keys=True))


def test_sceneio_symbols_exported(io):
    cdll = pathtracer_b200.load().cdll
    for name in _abi.SCENEIO_SYMBOLS:
        assert hasattr(cdll, "ptb_" + name), name
    hdr = open(os.path.join(os.path.dirname(GOLD), "..", "..", "include", "ptb_sceneio.h")).read()
    for name in _abi.SCENEIO_SYMBOLS:
        assert "ptb_" + name + "(" in hdr, f"{name} is bound but not declared in include/ptb_sceneio.h"


@pytest.mark.parametrize("name", sio.IMAGES)
def test_images_match_reference_decoder(io, gold, name):
    with sio.in_assets():
        assert np.array_equal(sio.dump_image(io, name), gold[0][f"image/{name}"])


@pytest.mark.parametrize("name,kind", sio.TEXTURES)
def test_texture_values_bit_exact(io, gold, name, kind):
    with sio.in_assets():
        got = sio.dump_texture(io, name, kind)
    want = gold[0][f"texture/{name}/{kind}"]
    assert got.dtype == want.dtype and np.array_equal(got.view(np.uint32) if not np.isnan(want).any() else np.nan_to_num(got, nan=-7),
                                                      want.view(np.uint32) if not np.isnan(want).any() else np.nan_to_num(want, nan=-7))


@pytest.mark.parametrize("name,lt", sio.MESHES)
def test_mesh_readers_match_reference(io, gold, name, lt):
    with sio.in_assets():
        d = sio.dump_mesh(io, name, lt)
    for key in ("vertices", "normals", "uvs", "vertex_colors", "tri"):
        assert np.array_equal(d[key], gold[0][f"mesh/{name}/{lt}/{key}"]), key
    meta = gold[1][f"mesh/{name}/{lt}"]
    assert d["n_groups"] == meta["n_groups"] and d["groups"] == meta["groups"]
    assert _norm({str(g): {str(k): v for k, v in per.items()} for g, per in d["slots"].items()}) == meta["slots"]


@pytest.mark.parametrize("name", sio.YARNS)
def test_yarn_reader_matches_reference(io, gold, name):
    """Yarns::Yarns(filename) (TriangleMesh.h:268-288): the same segments, bit for bit (as a set: the reference's constructor reorders them)."""
    with sio.in_assets():
        got = sio.dump_yarn(io, name)
    want = gold[0][f"yarn/{name}"]
    assert got.shape == want.shape == (16 + 1 + 4, 7) and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert (got[:, 6] == np.float32(0.1)).all()


def test_obj_reader_details(io):
    """What the fixture is built to exercise, stated explicitly (independent of the golden file)."""
    with sio.in_assets():
        d = sio.dump_mesh(io, "relief.obj", 1)
    tri = d["tri"]
    assert len(d["vertices"]) == 64 + 5 and len(tri) == 101                      # quads -> 2, two explicit triangles, pentagon -> 3
    assert (tri[:, :3] >= 0).all() and tri[:, :3].max() == 68                        # negative indices resolved
    assert set(np.unique(tri[:, 9])) == {0, 1}                                      # stone (0, re-used by name) and glass (1)
    assert (tri[-3:, 0] == 64).all() and (tri[-3:, 3:6] == -1).all() and (tri[-3:, 6:9] == 64).all()   # fan around its first vertex, v//vn form
    assert d["groups"]["glass"] == 1 and d["n_groups"] == 3                       # "never_used" inserted by the MTL pass, id 0
    kd0, kd1 = d["slots"][0][_abi.KIND_KD], d["slots"][1][_abi.KIND_KD]
    assert kd0[0] == "checker.png" and np.allclose(kd0[1], [.9, .1, .1])            # overwritten by the unused material: the reference's operator[] quirk
    assert kd1[0] == "" and np.allclose(kd1[1], [.2, .3, .4])                       # the indented "\tKd 9 9 9" is not a statement
    assert d["slots"][1][_abi.KIND_ALPHA][0] == "alpha.pgm" and d["slots"][0][_abi.KIND_NORMAL][0] == "bumps.bmp"
    assert np.allclose(d["slots"][1][_abi.KIND_NE][1], [12, 13, 14]) and np.allclose(d["slots"][0][_abi.KIND_NE][1], [40, 40, 40])
    assert d["slots"][0][_abi.KIND_KS][0] == "grey.png"


@pytest.mark.parametrize("name", sio.SCENES)
def test_scn_parser_matches_reference(io, gold, name):
    with sio.in_assets():
        d = sio.dump_scn(io, name)
    want = gold[1][f"scn/{name}"]
    got = _norm(d)
    assert got["header"] == want["header"]
    assert len(got["objects"]) == len(want["objects"])
    for i, (a, b) in enumerate(zip(got["objects"], want["objects"])):
        assert a == b, f"object {i}"


def test_scn_details_and_roundtrip(io, tmp_path):
    with sio.in_assets():
        d = sio.dump_scn(io, "full.scn")
        h = C.c_void_p()
        io.check(io.scn_load(b"full.scn", None, C.byref(h)))
        out = str(tmp_path / "resaved.scn")
        io.check(io.scn_save(h, out.encode()))
        io.scn_free(h)
        d2 = sio.dump_scn(io, out)
    m = d["objects"][3]
    assert m["xform"]["scale"] == 30.0 and m["xform"]["translation"] == [0.0, -17.0, 0.0], "keys at frames -5,-1: the last key places the object at frame 0"
    assert d["objects"][5]["xform"]["scale"] == 9.0, "keys at frames 3,7: the first key"
    assert d["objects"][1]["flip_normals"] == 1 and d["objects"][1]["is_envmap"] == 1
    assert d["objects"][3]["interp_normals"] == 1 and d["objects"][5]["interp_normals"] == 1, "TriMesh::init forces interp_normals"
    for a, b in zip(d["objects"], d2["objects"]):                                  # save -> load is the identity on everything but key frames
        a, b = dict(a), dict(b)
        a.pop("n_keyframes"); b.pop("n_keyframes")
        assert a == b
    assert {k: v for k, v in d["header"].items()} == d2["header"]


def test_reader_errors(io, tmp_path):
    h = C.c_void_p()
    assert io.scn_load(b"/nonexistent/x.scn", None, C.byref(h)) == -1 and b"cannot open" in io.sceneio_last_error()
    assert io.meshfile_read(b"/nonexistent/x.obj", 0, C.byref(h)) == -1
    assert io.meshfile_read(b"mesh.wrl", 0, C.byref(h)) == -5
    p, w, hh = C.POINTER(C.c_uint8)(), C.c_int32(), C.c_int32()
    jpg = tmp_path / "x.jpg"
    jpg.write_bytes(b"\xff\xd8\xff\xe0" + b"\0" * 32)
    assert io.image_load(str(jpg).encode(), C.byref(p), C.byref(w), C.byref(hh)) == -5 and b"JPEG" in io.sceneio_last_error()
    bad = tmp_path / "bad.scn"
    bad.write_text("W,H: 10, 10\nnrays: 1\nCam: broken\n")
    assert io.scn_load(str(bad).encode(), None, C.byref(h)) == -1 and b"Cam:" in io.sceneio_last_error()
    trunc = tmp_path / "t.png"
    trunc.write_bytes(open(os.path.join(sio.ASSETS, "checker.png"), "rb").read()[:60])
    assert io.image_load(str(trunc).encode(), C.byref(p), C.byref(w), C.byref(hh)) == -1


def test_malformed_files_are_errors_not_crashes(io, tmp_path):
    """Sizes and offsets taken from file headers are validated: an over-read, a gigabyte allocation or a C++ exception crossing the
    C boundary would take the host process down."""
    import struct
    p, w, hh = C.POINTER(C.c_uint8)(), C.c_int32(), C.c_int32()
    h = C.c_void_p()
    # 8-bpp BMP whose info-header size points the palette far outside the file
    px = bytes(range(16))
    bmp = b"BM" + struct.pack("<IHHI", 54 + len(px), 0, 0, 54) + struct.pack("<IiiHHIIiiII", 0x7fffff00, 4, 4, 1, 8, 0, len(px), 0, 0, 0, 0) + px
    f = tmp_path / "pal.bmp"; f.write_bytes(bmp)
    assert io.image_load(str(f).encode(), C.byref(p), C.byref(w), C.byref(hh)) == -1 and b"palette" in io.sceneio_last_error()
    # PNG header that announces 2^24 x 2^24 pixels
    png = open(os.path.join(sio.ASSETS, "checker.png"), "rb").read()
    import zlib
    ihdr = struct.pack(">IIBBBBB", 1 << 24, 1 << 24, 8, 2, 0, 0, 0)
    big = png[:8] + struct.pack(">I", 13) + b"IHDR" + ihdr + struct.pack(">I", zlib.crc32(b"IHDR" + ihdr)) + png[33:]
    f = tmp_path / "big.png"; f.write_bytes(big)
    assert io.image_load(str(f).encode(), C.byref(p), C.byref(w), C.byref(hh)) in (-1, -5)
    # OFF header that promises two billion vertices
    f = tmp_path / "big.off"; f.write_text("OFF\n2000000000 1 0\n0 0 0\n")
    assert io.meshfile_read(str(f).encode(), 0, C.byref(h)) == -1 and b"exceed" in io.sceneio_last_error()
    # OBJ face with relative references that point before the start of the vt / vn lists
    f = tmp_path / "neg.obj"
    f.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvn 0 0 1\nf -3/-5/-5 -2/-1/-1 -1/-2/-1\n")
    assert io.meshfile_read(str(f).encode(), 0, C.byref(h)) == -1 and b"before the start" in io.sceneio_last_error()
    f.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvn 0 0 1\nf -3/-2/-1 -2/-1/-1 -1/-2/-1\n")      # in range: fine
    assert io.meshfile_read(str(f).encode(), 0, C.byref(h)) == 0
    io.meshfile_free(h)
    # .yarn files whose counts the points do not honour, and one without a segment
    fp = C.POINTER(C.c_float)
    a, b, r, n = fp(), fp(), fp(), C.c_int32()
    f = tmp_path / "short.yarn"; f.write_text("2\n3\n0 0 0\n1 0 0\n1 1 0\n4\n0 0 1\n")
    assert io.yarnfile_read(str(f).encode(), C.byref(a), C.byref(b), C.byref(r), C.byref(n)) == -1 and b"fewer points" in io.sceneio_last_error()
    f.write_text("1\n1\n0 0 0\n")
    assert io.yarnfile_read(str(f).encode(), C.byref(a), C.byref(b), C.byref(r), C.byref(n)) == -1 and b"no segment" in io.sceneio_last_error()
    f.write_text("two yarns\n")
    assert io.yarnfile_read(str(f).encode(), C.byref(a), C.byref(b), C.byref(r), C.byref(n)) == -1
    assert io.yarnfile_read(str(tmp_path / "absent.yarn").encode(), C.byref(a), C.byref(b), C.byref(r), C.byref(n)) == -1


def _scn_with_modes(tmp_path):
    """old.scn with a participating medium, a ghost ground plane and a background photograph switched on."""
    txt = open(os.path.join(sio.ASSETS, "old.scn")).read()
    assert "fog_density: 0.000000\nfog_type: 0" in txt and "nbobjects: 4" in txt
    txt = txt.replace("fog_density: 0.000000\nfog_type: 0", "fog_density: 0.250000\nfog_absorption: 0.200000\nfog_density_decay: 0.050000\nfog_absorption_decay: 0.050000\n"
                      "fog_type: 1\nfog_phase_type: 2\ndouble_frustum_start_t: 0.000000")
    txt = txt.replace("nbobjects: 4", "background: checker.png\nnbobjects: 4")
    txt = txt.replace("name: Plane\nmiroir: 0\n", "name: Plane\nmiroir: 0\nghost: 1\n")
    f = tmp_path / "modes.scn"
    f.write_text(txt)
    return str(f)


def test_scn_fog_ghost_background_reach_the_renderer(port, ref, tmp_path):
    """A .scn that switches on fog, a ghost object and a background photograph: the reference's own load_scene + render against
    the oracle fed by the PRODUCT's reader (Python mirror and the C-level ptb_load_scene route), bit for bit."""
    f = _scn_with_modes(tmp_path)
    with sio.in_assets():
        RIO = sio.sceneio_of(ref.cdll, "ref_")
        ctx = C.c_void_p()
        ref.check(ref.create(0, C.byref(ctx)))
        cam, p = _abi.Camera(), _abi.Params()
        RIO.check(RIO.load_scene(ctx, f.encode(), None, C.byref(cam), C.byref(p)))
        ref.check(ref.commit(ctx), ctx)
        ref.check(ref.set_option(ctx, _abi.ORC_OPT_THREADS, 1), ctx)
        want, st = np.empty((p.H, p.W, 3), np.float32), _abi.Stats()
        ref.check(ref.render(ctx, C.byref(cam), C.byref(p), _abi.fptr(want), None, None, C.byref(st)), ctx)
        ref.destroy(ctx)
        rt = api.Raytracer(port).load_scene(f)
        assert rt.s.fog_density == pytest.approx(0.25) and rt.s.objects[2].ghost and rt.s.background is not None
        rt.commit()
        rt.set_option(_abi.ORC_OPT_THREADS, 1)
        got = rt.render_image_nopreviz().copy()
    assert np.array_equal(got, want)
    assert [rt.stats["rays_closest"], rt.stats["rays_shadow"]] == [st.rays_closest, st.rays_shadow]


@pytest.mark.parametrize("frame", [0, 3])
def test_scn_keyframes_straddling_the_frame_are_slerped(port, ref, io, tmp_path, frame):
    """full.scn with its keys moved to frames -5 / 4 and -1 / 7: at frames 0 and 3 the placement is an interpolation (Slerp for the
    rotation).  Reference load_scene + render at Scene::current_frame == oracle fed by the product reader, bit for bit; and the
    parsed keys themselves equal the reference's maps."""
    txt = open(os.path.join(sio.ASSETS, "full.scn")).read()
    for old, new in (("\n-1.000000 ", "\n4.000000 "), ("\n3.000000 ", "\n-1.000000 ")):
        assert old in txt
        txt = txt.replace(old, new)
    f = tmp_path / "keys.scn"
    f.write_text(txt)
    with sio.in_assets():
        RIO = sio.sceneio_of(ref.cdll, "ref_")
        ctx = C.c_void_p()
        ref.check(ref.create(0, C.byref(ctx)))
        cam, p = _abi.Camera(), _abi.Params()
        RIO.check(RIO.load_scene(ctx, str(f).encode(), None, C.byref(cam), C.byref(p)))
        ref.check(ref.set_frame(ctx, float(frame)), ctx)
        ref.check(ref.commit(ctx), ctx)
        ref.check(ref.set_option(ctx, _abi.ORC_OPT_THREADS, 1), ctx)
        want = np.empty((p.H, p.W, 3), np.float32)
        ref.check(ref.render(ctx, C.byref(cam), C.byref(p), _abi.fptr(want), None, None, None), ctx)
        h = C.c_void_p()
        io.check(io.scn_load(str(f).encode(), None, C.byref(h)))
        for obj in range(6):
            for kind, width in ((_abi.KEY_SCALE, 1), (_abi.KEY_TRANSLATION, 3), (_abi.KEY_ROTATION, 9)):
                n = io.scn_get_keyframes(h, obj, kind, None, None, 0)
                assert n == RIO.scn_get_keyframes(ctx, obj, kind, None, None, 0)
                a, b = np.zeros((2, max(n, 1)), np.float32), np.zeros((2, max(n, 1) * width), np.float32)
                io.scn_get_keyframes(h, obj, kind, _abi.fptr(a[0]), _abi.fptr(b[0]), n)
                RIO.scn_get_keyframes(ctx, obj, kind, _abi.fptr(a[1]), _abi.fptr(b[1]), n)
                assert np.array_equal(a[0], a[1]) and np.array_equal(b[0], b[1]), (obj, kind)
        io.scn_free(h)
        ref.destroy(ctx)
        rt = api.Raytracer(port).load_scene(str(f))
        assert sum(len(o.rotation_keyframes) for o in rt.s.objects) == 4
        rt.s.current_frame = frame
        rt.commit()
        rt.set_option(_abi.ORC_OPT_THREADS, 1)
        got = rt.render_image_nopreviz().copy()
    assert np.array_equal(got, want)


def test_unsupported_scene_features_are_refused(port, tmp_path):
    txt = open(os.path.join(sio.ASSETS, "old.scn")).read()
    with sio.in_assets():
        for bad, what in ((txt.replace("nb_textures: 0\nnb_normalmaps: 0", "nb_textures: 0\nnb_normalmaps: 0\nnb_subsurfaces: 1\ntexture: Color: (128.000000, 100.000000, 80.000000)\nmultiplier: (0.500000, 0.400000, 0.300000)", 1), "subsurface"),):
            f = tmp_path / f"{what}.scn"
            f.write_text(bad)
            try:
                rt = api.Raytracer(port).load_scene(str(f))
            except _abi.PtbError:
                continue                      # the reader itself refused the block
            with pytest.raises(_abi.PtbError, match="unsupported"):
                rt.commit()


@pytest.mark.parametrize("name", sio.RENDER_SCENES)
def test_oracle_fed_by_product_reader_equals_reference_own_load(port, gold, name):
    """The reference's `load_scene` + render (committed golden) == oracle/port rendering what the PRODUCT's reader produced: bit for bit."""
    g = gold[0]
    with sio.in_assets():
        rt = api.Raytracer(port).load_scene(name).commit()
    rt.set_option(_abi.ORC_OPT_THREADS, 1)
    obj, tri, t = rt.primary_ids()
    assert np.array_equal(obj, g[f"render/{name}/obj"]) and np.array_equal(tri, g[f"render/{name}/tri"]) and np.array_equal(t, g[f"render/{name}/t"])
    img = rt.render_image_nopreviz()
    assert np.array_equal(img, g[f"render/{name}/imagedouble"]) and np.array_equal(rt.sample_count, g[f"render/{name}/sample_count"])
    assert np.array_equal(rt.image, g[f"render/{name}/image"])
    assert [rt.stats["rays_closest"], rt.stats["rays_shadow"]] == g[f"render/{name}/rays"].tolist()


def test_live_reference_readers_agree(io, ref):
    """When oracle/_ref is here: the same walk over the reference's own readers, live."""
    rio = sio.sceneio_of(ref.cdll, "ref_")
    with sio.in_assets():
        for n in sio.IMAGES:
            assert np.array_equal(sio.dump_image(io, n), sio.dump_image(rio, n)), n
        for n, lt in sio.MESHES:
            a, b = sio.dump_mesh(io, n, lt), sio.dump_mesh(rio, n, lt)
            assert all(np.array_equal(a[k], b[k]) if isinstance(a[k], np.ndarray) else a[k] == b[k] for k in a), n
        for n in sio.SCENES:
            assert sio.dump_scn(io, n) == sio.dump_scn(rio, n), n
        for n in sio.YARNS:
            assert np.array_equal(sio.dump_yarn(io, n), sio.dump_yarn(rio, n)), n


# ---- GPU: the product end to end from a file ------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", sio.RENDER_SCENES)
def test_gpu_renders_scene_files(gpu, port, gold, name):
    g = gold[0]
    with sio.in_assets():
        a = api.Raytracer(gpu).load_scene(name).commit()            # Python mirror of Raytracer::load_scene
        b = api.Raytracer(gpu).load_scene_native(name)              # ptb_load_scene, the C route
    for rt in (a, b):
        obj, tri, t = rt.primary_ids()
        same = (obj == g[f"render/{name}/obj"]) & (tri == g[f"render/{name}/tri"])
        assert same.mean() >= 0.999, "primary ids against the reference's own load + picking query"
        img = rt.render_image_nopreviz().copy()
        check_images(img, g[f"render/{name}/imagedouble"], frac=0.01)
        assert np.allclose(rt.sample_count, g[f"render/{name}/sample_count"], rtol=1e-5)
    ia, ib = a.render_image_nopreviz().copy(), b.render_image_nopreviz().copy()
    assert np.allclose(ia, ib, rtol=1e-5), "both load routes build the same scene"
