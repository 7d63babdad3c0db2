"""Shared by tests/test_scene_io.py and tests/golden/make_golden.py: walk an implementation of the include/ptb_sceneio.h accessors
(the product's readers in libptb200.so, or the reference's own readers exported by oracle/_ref with the prefix ref_) into plain
numpy arrays / dicts, so that one comparison covers both."""
import contextlib
import ctypes as C
import os

import numpy as np

from pathtracer_b200 import _abi

ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assets")
IMAGES = ["checker.png", "grey.png", "pal.png", "rgba.png", "grey16.png", "bumps.bmp", "bumps8.bmp", "alpha.pgm", "tint.ppm", "sky.tga", "sky_rle.tga", "grey.tga"]
TEXTURES = [("checker.png", 0), ("bumps.bmp", 1), ("alpha.pgm", 0), ("grey16.png", 1)]
MESHES = [("relief.obj", 0), ("relief.obj", 1), ("sheet.obj", 1), ("tetra.obj", 1), ("octa.off", 0)]
YARNS = ["weave.yarn"]
SCENES = ["full.scn", "forms.scn", "old.scn"]
RENDER_SCENES = ["full.scn", "old.scn"]


@contextlib.contextmanager
def in_assets():
    """The reference opens the files a .scn names relative to the working directory."""
    old = os.getcwd()
    os.chdir(ASSETS)
    try:
        yield
    finally:
        os.chdir(old)


def sceneio_of(lib_cdll, prefix):
    return _abi.SceneIO(lib_cdll, prefix)


def dump_image(io, name):
    p, w, h = C.POINTER(C.c_uint8)(), C.c_int32(), C.c_int32()
    io.check(io.image_load(name.encode(), C.byref(p), C.byref(w), C.byref(h)))
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()
    io.image_free(p)
    return a


def dump_texture(io, name, kind):
    p, w, h = C.POINTER(C.c_float)(), C.c_int32(), C.c_int32()
    io.check(io.texture_load(name.encode(), kind, C.byref(p), C.byref(w), C.byref(h)))
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()
    io.image_free(p)
    return a


def dump_yarn(io, name):
    """-> (n, 7) rows A, B, R sorted lexicographically (the reference hands its segments back in the order its BVH build left)"""
    fp = C.POINTER(C.c_float)
    a, b, r, n = fp(), fp(), fp(), C.c_int32()
    io.check(io.yarnfile_read(name.encode(), C.byref(a), C.byref(b), C.byref(r), C.byref(n)))
    rows = np.concatenate([np.ctypeslib.as_array(a, shape=(n.value, 3)), np.ctypeslib.as_array(b, shape=(n.value, 3)),
                           np.ctypeslib.as_array(r, shape=(n.value, 1))], 1).astype(np.float32, copy=True)
    for p in (a, b, r):
        io.yarnfile_free(p)
    return rows[np.lexsort(rows.T[::-1])]


def _slot(sl):
    return os.path.basename(sl.file.decode()), [float(x) for x in sl.mult]


def dump_mesh(io, name, load_textures):
    """-> dict of arrays + {"groups": {name: id}, "slots": {group: {kind: (file basename, mult)}}}"""
    h = C.c_void_p()
    io.check(io.meshfile_read(name.encode(), load_textures, C.byref(h)))
    info = _abi.MeshfileInfo()
    io.check(io.meshfile_get(h, C.byref(info)))
    arr = lambda p, n, k, dt: (np.ctypeslib.as_array(p, shape=(n, k)).astype(dt, copy=True) if n else np.zeros((0, k), dt))
    d = {"vertices": arr(info.vertices, info.n_vertices, 3, np.float32), "normals": arr(info.normals, info.n_normals, 3, np.float32),
         "uvs": arr(info.uvs, info.n_uvs, 2, np.float32), "vertex_colors": arr(info.vertex_colors, info.n_vertex_colors, 3, np.float32),
         "tri": arr(info.tri, info.n_tri, 10, np.int32), "n_groups": info.n_groups, "groups": {}, "slots": {}}
    for g in range(info.n_groups):
        buf = C.create_string_buffer(_abi.PATH_MAX)
        if io.meshfile_group_name(h, g, buf) == _abi.OK:
            d["groups"][buf.value.decode()] = g
    if load_textures:
        g = 0
        while True:
            sl, per = _abi.Slot(), {}
            for kind in range(_abi.N_KINDS):
                if io.meshfile_group_slot(h, g, kind, C.byref(sl)) == _abi.OK:
                    per[kind] = _slot(sl)
            if not per:
                break
            d["slots"][g] = per
            g += 1
    io.meshfile_free(h)
    return d


def _fields(st, skip=()):
    out = {}
    for name, typ in st._fields_:
        if name in skip:
            continue
        v = getattr(st, name)
        if isinstance(v, bytes):
            out[name] = os.path.basename(v.decode())
        elif isinstance(v, C.Structure):
            out[name] = _fields(v)
        elif hasattr(v, "__len__"):
            out[name] = [float(x) if isinstance(x, float) else int(x) for x in v]
        else:
            out[name] = float(v) if isinstance(v, float) else int(v)
    return out


def dump_scn(io, name):
    """-> {"header": {...}, "objects": [{fields..., "slots": {kind: [(file, mult), ...]}}]}"""
    h = C.c_void_p()
    io.check(io.scn_load(name.encode(), None, C.byref(h)))
    hd = _abi.ScnHeader()
    io.check(io.scn_get_header(h, C.byref(hd)))
    d = {"header": _fields(hd), "objects": []}
    for i in range(hd.n_objects):
        o = _abi.ScnObject()
        io.check(io.scn_get_object(h, i, C.byref(o)))
        # fields the reference only defines for the object's own type stay out of the comparison
        skip = {"csv_file"}
        if o.type != _abi.SCN_SPHERE:
            skip |= {"is_envmap", "envmap", "O", "R"}
        if o.type != _abi.SCN_PLANE:
            skip |= {"A", "N"}
        if o.type != _abi.SCN_MESH:
            skip |= {"is_centered", "has_csv"}
        od = _fields(o, skip)
        od["slots"] = {}
        for kind in range(_abi.N_KINDS):
            od["slots"][kind] = []
            for g in range(o.n_slots[kind]):
                sl = _abi.Slot()
                io.check(io.scn_get_slot(h, i, kind, g, C.byref(sl)))
                od["slots"][kind].append(_slot(sl))
        d["objects"].append(od)
    io.scn_free(h)
    return d
