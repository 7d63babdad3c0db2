#!/usr/bin/env python3
"""Write the small scene-file fixtures under tests/golden/assets/ (committed): meshes, images and .scn files that exercise
every branch of the readers (face formats, fans, negative indices, groups, MTL statements, optional .scn lines).
Needs PIL for the image encoders; the tests only read the committed files.   python tests/golden/make_assets.py"""
import math
import os

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "assets")


def images():
    y, x = np.mgrid[0:24, 0:32]
    chk = np.zeros((24, 32, 3), np.uint8)
    chk[..., 0] = 40 + 180 * (((x // 4) + (y // 4)) % 2)
    chk[..., 1] = (x * 8) % 256
    chk[..., 2] = (y * 10) % 256
    Image.fromarray(chk).save(os.path.join(OUT, "checker.png"))
    Image.fromarray(chk[..., 1]).save(os.path.join(OUT, "grey.png"))                         # 8-bit grey
    Image.fromarray(chk).convert("P", palette=Image.ADAPTIVE, colors=16).save(os.path.join(OUT, "pal.png"))   # palette
    rgba = np.dstack([chk, 255 - chk[..., :1]])
    Image.fromarray(rgba, "RGBA").save(os.path.join(OUT, "rgba.png"))                         # alpha dropped by the reader
    g16 = (chk[..., 1].astype(np.uint16) * 257) ^ 0x55
    Image.fromarray(g16, "I;16").save(os.path.join(OUT, "grey16.png"))                        # 16-bit: high byte kept
    nrm = np.zeros((16, 16, 3), np.uint8)
    yy, xx = np.mgrid[0:16, 0:16]
    nrm[..., 0] = (128 + 60 * np.sin(xx * math.pi / 4)).astype(np.uint8)
    nrm[..., 1] = (128 + 60 * np.cos(yy * math.pi / 4)).astype(np.uint8)
    nrm[..., 2] = 230
    Image.fromarray(nrm).save(os.path.join(OUT, "bumps.bmp"))                                 # 24-bit BMP, bottom-up
    Image.fromarray(nrm).convert("P", palette=Image.ADAPTIVE, colors=32).save(os.path.join(OUT, "bumps8.bmp"))   # 8-bit palette BMP
    a = (255 * (((xx // 2) + (yy // 2)) % 4 != 0)).astype(np.uint8)
    with open(os.path.join(OUT, "alpha.pgm"), "wb") as f:
        f.write(b"P5\n# holes\n16 16\n255\n" + a.tobytes())
    with open(os.path.join(OUT, "tint.ppm"), "wb") as f:
        f.write(b"P6 16 16 255\n" + nrm.tobytes())
    sky = np.zeros((16, 32, 3), np.uint8)
    sy, sx = np.mgrid[0:16, 0:32]
    sky[..., 0] = 60 + 6 * sy
    sky[..., 1] = 90 + 5 * sy
    sky[..., 2] = 200 - 3 * sx
    sky[3:5, 20:23] = 255
    Image.fromarray(sky).save(os.path.join(OUT, "sky.tga"))                                   # uncompressed, origin per PIL default
    Image.fromarray(sky).save(os.path.join(OUT, "sky_rle.tga"), compression="tga_rle")
    Image.fromarray(sky[..., 0]).save(os.path.join(OUT, "grey.tga"))


def meshes():
    n = 7                                       # (n+1)^2 vertices of a bumpy sheet, quads as 4-vertex faces (fans)
    vs, vts, vns = [], [], []
    for j in range(n + 1):
        for i in range(n + 1):
            u, v = i / n, j / n
            h = 0.25 * math.sin(3.1 * u * math.pi) * math.cos(2.3 * v * math.pi)
            vs.append((2 * u - 1, h, 2 * v - 1))
            dx = 0.25 * 3.1 * math.pi * math.cos(3.1 * u * math.pi) * math.cos(2.3 * v * math.pi) / 2
            dz = -0.25 * 2.3 * math.pi * math.sin(3.1 * u * math.pi) * math.sin(2.3 * v * math.pi) / 2
            nn = np.array([-dx, 1, -dz]); nn /= np.linalg.norm(nn)
            vns.append(tuple(nn)); vts.append((u * 2.5, v * 1.5))
    idx = lambda i, j: j * (n + 1) + i + 1
    L = ["# bumpy sheet: every face statement form the reader knows", "mtllib relief.mtl", "o sheet"]
    L += [f"v {x:.6f} {y:.6f} {z:.6f}" for x, y, z in vs]
    L += [f"vt {u:.6f} {v:.6f}" for u, v in vts]
    L += [f"vn {x:.6f} {y:.6f} {z:.6f}" for x, y, z in vns]
    total = len(vs)
    for j in range(n):
        if j == 0:
            L.append("usemtl stone")
        if j == 3:
            L.append("usemtl glass")
        if j == 5:
            L.append("usemtl stone")             # a repeated name keeps its first id
        for i in range(n):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            form = (i + j) % 5
            if form == 0:
                L.append(f"f {a}/{a}/{a} {d}/{d}/{d} {c}/{c}/{c} {b}/{b}/{b}")                       # quad, v/vt/vn
            elif form == 1:
                L.append(f"f {a}//{a} {d}//{d} {c}//{c}"); L.append(f"f {a}//{a} {c}//{c} {b}//{b}")    # v//vn triangles
            elif form == 2:
                L.append(f"f {a}/{a} {d}/{d} {c}/{c} {b}/{b}")                                       # quad, v/vt
            elif form == 3:
                L.append(f"f {a - total - 1}/{a - total - 1}/{a - total - 1} {d - total - 1}/{d - total - 1}/{d - total - 1} "
                         f"{c - total - 1}/{c - total - 1}/{c - total - 1} {b - total - 1}/{b - total - 1}/{b - total - 1}")   # negative indices
            else:
                L.append(f"f {a} {d} {c} {b} ")                                                     # bare indices, trailing blank
    # the render-safe variant: the same sheet with v/vt/vn on EVERY face (the reference shades faces without normals from
    # uninitialised memory, SURVEY.md App. D#12, and so cannot pin anything for them)
    # uv inside [0,1): the reference's normal-map lookup does not wrap (TriangleMesh.cpp:961) and reads out of bounds beyond it
    S = [l for l in L if not l.startswith(("f ", "usemtl", "vt "))] + [f"vt {u / 2.5 * 0.98 + 0.01:.6f} {v / 1.5 * 0.98 + 0.01:.6f}" for u, v in vts]
    for j in range(n):
        if j in (0, 5):
            S.append("usemtl stone")
        if j == 3:
            S.append("usemtl glass")
        for i in range(n):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            r = lambda k: f"{k}/{k}/{k}"
            m = lambda k: f"{k - total - 1}/{k - total - 1}/{k - total - 1}"
            if (i + j) % 3 == 0:
                S.append(f"f {r(a)} {r(d)} {r(c)} {r(b)}")
            elif (i + j) % 3 == 1:
                S.append(f"f {m(a)} {m(d)} {m(c)} {m(b)} ")
            else:
                S.append(f"f {r(a)} {r(d)} {r(c)}"); S.append(f"f {r(a)} {r(c)} {r(b)}")
    open(os.path.join(OUT, "sheet.obj"), "w").write("\n".join(S).replace("relief.mtl", "sheet.mtl") + "\n")
    open(os.path.join(OUT, "sheet.mtl"), "w").write("newmtl stone\nKd 0.7 0.6 0.5\nmap_Kd checker.png\nnewmtl glass\nKd 0.2 0.3 0.4\n")
    # a pentagon fan on top, still in group stone
    base = len(vs)
    for k in range(5):
        ang = 2 * math.pi * k / 5
        L.append(f"v {0.3 * math.cos(ang):.6f} 0.600000 {0.3 * math.sin(ang):.6f}")
    L.append("vn 0 1 0")
    vn_last = len(vns) + 1
    L.append("f " + " ".join(f"{base + 1 + k}//{vn_last}" for k in range(5)))
    open(os.path.join(OUT, "relief.obj"), "w").write("\n".join(L) + "\n")
    open(os.path.join(OUT, "relief.mtl"), "w").write(
        "# materials\nnewmtl stone\nKd 0.700000 0.600000 0.500000\nKs 0.100000 0.100000 0.100000\nNs 40.000000\nillum 2\nmap_Kd checker.png\nmap_Bump bumps.bmp\n"
        "\nnewmtl glass\nKd 0.2 0.3 0.4\nKs 0.5 0.4 0.3\nNs 12 13 14\nmap_d alpha.pgm\n\tKd 9 9 9\n"
        "\nnewmtl never_used\nKd 0.9 0.1 0.1\nmap_Ks grey.png\n")
    # an OBJ without usemtl / mtllib: one "Default" group
    open(os.path.join(OUT, "tetra.obj"), "w").write(
        "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 -1\nvn 0 -1 0\nvn -1 0 0\nvn 0.577 0.577 0.577\n"
        "f 1//1 3//1 2//1\nf 1//2 2//2 4//2\nf 1//3 4//3 3//3\nf 2//4 3//4 4//4\n")
    open(os.path.join(OUT, "octa.off"), "w").write(
        "OFF\n6 8 12\n1 0 0\n-1 0 0\n0 1 0\n0 -1 0\n0 0 1\n0 0 -1\n"
        "3 0 2 4\n3 2 1 4\n3 1 3 4\n3 3 0 4\n3 2 0 5\n3 1 2 5\n3 3 1 5\n3 0 3 5\n")


OBJ_TAIL = """nb_textures: {nt}
{textures}nb_normalmaps: {nn}
{normals}{subs}nb_specularmaps: {ns}
{speculars}nb_alphamaps: {na}
{alphas}nb_expmaps: {ne}
{exps}nb_transpmaps: {ntr}
{transps}nb_refrindexmaps: {nr}
{refrs}"""


def slots(entries, single=False):
    out = ""
    for name, mult in entries:
        out += f"texture: {name}\n"
        out += f"multiplier: {mult[0]:f})\n" if single else f"multiplier: ({mult[0]:f}, {mult[1]:f}, {mult[2]:f})\n"
    return out


def obj_block(head, name, *, miroir=0, ghost=None, tr=(0, 0, 0), rot=(1, 0, 0, 0, 1, 0, 0, 0, 1), center=(0, 0, 0), scale=1.0, interp=1, flip=0, keys=None,
              tex=(), nrm=(), sub=None, spec=(), alpha=(), exp=(), transp=(), refr=(), tail=""):
    s = f"{head}\nname: {name}\nmiroir: {miroir}\n"
    if ghost is not None:
        s += f"ghost: {ghost}\n"
    s += f"translation: ({tr[0]:f}, {tr[1]:f}, {tr[2]:f})\n"
    s += "rotation: (" + ", ".join(f"{v:f}" for v in rot) + ")\n"
    s += f"center: ({center[0]:f}, {center[1]:f}, {center[2]:f})\nscale: {scale:f}\ndisplay_edges: 0\ninterp_normals: {interp}\nflip_normals: {flip}\n"
    if keys is not None:
        s += f"nb_transforms: {len(keys)}\n"
        s += "".join(f"{k[0]:f} {k[1]:f}\n" for k in keys)
        s += "".join(f"{k[0]:f} {k[2][0]:f}, {k[2][1]:f}, {k[2][2]:f}\n" for k in keys)
        s += "".join(f"{k[0]:f} " + ", ".join(f"{v:f}" for v in k[3]) + "\n" for k in keys)
    subs = "" if sub is None else f"nb_subsurfaces: {len(sub)}\n" + slots(sub)
    s += OBJ_TAIL.format(nt=len(tex), textures=slots(tex), nn=len(nrm), normals=slots(nrm), subs=subs, ns=len(spec), speculars=slots(spec),
                         na=len(alpha), alphas=slots(alpha), ne=len(exp), exps=slots(exp), ntr=len(transp), transps=slots(transp, True),
                         nr=len(refr), refrs=slots(refr, True))
    return s + tail


def scenes():
    full_scene("full.scn", "sheet.obj", "tetra.obj")       # renderable by the reference
    full_scene("forms.scn", "relief.obj", "octa.off")      # parse-level fixture: meshes the reference cannot shade
    old_scene()


def full_scene(out_name, MESH_A, MESH_B):
    lum = 1000000000 * 4. * math.pi / (4. * math.pi * 10 * 10 * math.pi)
    c, s_ = math.cos(0.4), math.sin(0.4)
    roty = (c, 0, s_, 0, 1, 0, -s_, 0, c)
    light = dict(tail="is_envmap: 0\nenvmapfilename: \nO: (10.000000, 23.000000, 15.000000)\nR: 10.000000\n")
    # ---- current format: every optional line present
    objs = [
        obj_block("NEW SPHERE", "Sphere", ghost=0, center=(10, 23, 15), keys=[], sub=[], **light),
        obj_block("NEW SPHERE", "Sphere", ghost=0, flip=1, keys=[], sub=[], tail="is_envmap: 1\nenvmapfilename: sky.tga\nO: (0.000000, 0.000000, 0.000000)\nR: 1000000.000000\n"),
        obj_block("NEW PLANE", "Plane", ghost=0, tr=(0, -27.3, 0), keys=[], sub=[], tex=[("Color: (200.000000, 180.000000, 160.000000)", (0.8, 0.7, 0.6))],
                  spec=[("Null", (0.1, 0.1, 0.1))], exp=[("Color: (30.000000, 30.000000, 30.000000)", (30, 30, 30))],
                  tail="Point: (0.000000, 0.000000, 0.000000)\nN: (0.000000, 1.000000, 0.000000)\n"),
        obj_block("NEW MESH", MESH_A, ghost=0, tr=(9, 9, 9), center=(0.1, 0.2, 0.3), scale=2.0,      # static placement overridden by the keys (all <= frame 0: last key)
                  keys=[(-5.0, 11.0, (1, 2, 3), (1, 0, 0, 0, 1, 0, 0, 0, 1)), (-1.0, 30.0, (0, -17, 0), roty)], sub=[("Null", (0, 0, 0))],
                  tex=[("checker.png", (0.9, 0.8, 0.7)), ("Color: (128.000000, 128.000000, 255.000000)", (0.5, 0.5, 1.0))],
                  nrm=[("bumps.bmp", (0, 0, 1)), ("Null", (0, 0, 1))], spec=[("Null", (0.2, 0.2, 0.2)), ("grey.png", (0.5, 0.5, 0.5))],
                  alpha=[("1.000000", (1, 1, 1)), ("alpha.pgm", (1, 1, 1))], exp=[("Null", (50, 50, 50)), ("Null", (20, 20, 20))],
                  transp=[("Null", (1.0,)), ("Null", (0.0,))], refr=[("Null", (1.3,)), ("Null", (1.5,))],
                  tail="is_centered: 1\nhas_csv: 0\ncsv_file: \n"),
        obj_block("NEW SPHERE", "Sphere", miroir=1, ghost=0, center=(-14, -20.3, 8), keys=[], sub=[],
                  tail="is_envmap: 0\nenvmapfilename: \nO: (-14.000000, -20.300000, 8.000000)\nR: 7.000000\n"),
        obj_block("NEW MESH", MESH_B, ghost=0, tr=(1, 1, 1), scale=1.0, keys=[(3.0, 9.0, (16, -21, 6), (1, 0, 0, 0, 1, 0, 0, 0, 1)), (7.0, 2.0, (0, 0, 0), roty)],
                  sub=[], interp=0, tail="is_centered: 0\nhas_csv: 0\ncsv_file: \n"),
    ]
    head = ("W,H: 80, 64\nnrays: 4\nnbframes: 1\nCam: (0.000000, 0.000000, 50.000000), (0.000000, -0.374607, -0.927184), (0.000000, 0.927184, -0.374607)\n"
            "fov: 0.610865\nfocus: 50.000000\naperture: 0.100000\nsigma_filter: 0.500000\ngamma: 2.200000\n"
            "is_lenticular: 0\nlenticular_nb_images: 10\nlenticular_max_angle: 0.261799\nlenticular_pixel_width: 10\nisArray: 0\nnbviewX: 1\nnbviewY: 1\n"
            "maxSpacingX: 1.000000\nmaxSpacingY: 1.000000\nbounces: 4\nhas_denoiser: 0\n"
            f"intensite_lum: {lum:f}\nintensite_envmap: 0.700000\nnbobjects: {len(objs)}\n")
    fog = ("fog_density: 0.000000\nfog_absorption: 0.000000\nfog_density_decay: 0.000000\nfog_absorption_decay: 0.000000\nfog_type: 0\nfog_phase_type: 0\n"
           "double_frustum_start_t: 0.000000\n")
    open(os.path.join(OUT, out_name), "w").write(head + "".join(objs) + fog)


def old_scene():
    lum = 1000000000 * 4. * math.pi / (4. * math.pi * 10 * 10 * math.pi)
    light = dict(tail="is_envmap: 0\nenvmapfilename: \nO: (10.000000, 23.000000, 15.000000)\nR: 10.000000\n")
    # ---- oldest format: no nbframes / lenticular / has_denoiser / ghost / nb_transforms / nb_subsurfaces / is_centered lines, short fog block
    objs = [
        obj_block("NEW SPHERE", "Sphere", center=(10, 23, 15), **light),
        obj_block("NEW SPHERE", "Sphere", flip=1, tail="is_envmap: 0\nenvmapfilename: \nO: (0.000000, 0.000000, 0.000000)\nR: 1000000.000000\n"),
        obj_block("NEW PLANE", "Plane", tr=(0, -27.3, 0), tail="Point: (0.000000, 0.000000, 0.000000)\nN: (0.000000, 1.000000, 0.000000)\n"),
        obj_block("NEW MESH", "tetra.obj", tr=(0, -20, 0), scale=14.0, tex=[("Color: (220.000000, 60.000000, 60.000000)", (0.86, 0.24, 0.24))],
                  tail="has_csv: 0\ncsv_file: \n"),
    ]
    head = ("W,H: 72, 48\nnrays: 3\nCam: (0.000000, 0.000000, 50.000000), (0.000000, -0.374607, -0.927184), (0.000000, 0.927184, -0.374607)\n"
            "fov: 0.610865\nfocus: 50.000000\naperture: 0.100000\nsigma_filter: 0.500000\ngamma: 2.200000\nbounces: 3\n"
            f"intensite_lum: {lum:f}\nintensite_envmap: 1.000000\nnbobjects: {len(objs)}\n")
    open(os.path.join(OUT, "old.scn"), "w").write(head + "".join(objs) + "fog_density: 0.000000\nfog_type: 0\n")


def yarns():
    """A .yarn file as Yarns::Yarns(filename) reads it (TriangleMesh.h:268-288): yarn count, then per yarn a point count and the
    points (times 50 in the reader).  Three yarns: a helix, a straight two-point one and a zig-zag; mixed separators and exponents."""
    lines = ["3"]
    t = np.linspace(0, 4 * math.pi, 17)
    lines.append(str(len(t)))
    lines += [f"{0.1 * math.cos(a):.6f} {0.02 * a - 0.5:.6f} {0.1 * math.sin(a):.6f}" for a in t]      # (y = -0.5 .. -0.25: in the default view)
    lines += ["2", "-2.5e-1 -0.4 1E-1", "0.25\t-0.35   -0.1"]
    lines.append("5")
    lines += [f"{-0.2 + 0.1 * k} {-0.3 + 0.05 * (k % 2)} {0.02 * k}" for k in range(5)]
    open(os.path.join(OUT, "weave.yarn"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--yarns-only" in __import__("sys").argv:
        yarns()
        raise SystemExit(0)
    images()
    meshes()
    scenes()
    yarns()
    print(sorted(os.listdir(OUT)))
