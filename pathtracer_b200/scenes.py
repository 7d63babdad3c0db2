"""Closed-form synthetic inputs for the benchmark configurations (BASELINE.json).

No RNG anywhere, so every implementation (CUDA path, oracle/_ref, oracle/port) regenerates
bit-identical arrays from the same parameters.  Meshes are handed over as a file reader would have
produced them (file axes, file order); TriMesh::init semantics (axis swap, centre/normalise) are the
receiving implementation's job.
"""
import math

import numpy as np

from .api import Cylinder, Plane, PointSet, Raytracer, Sphere, Texture, TriMesh, Yarns


def displaced_torus(nv):
    """4*nv^2 triangles.  Closed genus-1 surface without poles (a UV sphere's sliver pole triangles
    produce NaN barycentrics that crash the reference's alpha lookup)."""
    nu = 2 * nv
    i = np.arange(nu + 1, dtype=np.float64)[None, :]
    j = np.arange(nv + 1, dtype=np.float64)[:, None]
    u = 2 * math.pi * i / nu
    v = 2 * math.pi * j / nv
    n = np.stack([np.cos(v) * np.cos(u), np.sin(v) * np.ones_like(u), np.cos(v) * np.sin(u)], -1)
    rho = 0.4 * (1 + 0.15 * np.sin(9 * u) * np.sin(7 * v) + 0.04 * np.sin(40 * u + 3) * np.sin(33 * v))
    ring = np.stack([np.cos(u) * np.ones_like(v), np.zeros_like(u * v), np.sin(u) * np.ones_like(v)], -1)
    p = ring + rho[..., None] * n
    uv = np.stack([(i / nu) * np.ones_like(j), (j / nv) * np.ones_like(i)], -1)
    vertices = p.reshape(-1, 3).astype(np.float32)
    normals = n.reshape(-1, 3).astype(np.float32)
    uvs = uv.reshape(-1, 2).astype(np.float32)
    jj, ii = np.meshgrid(np.arange(nv), np.arange(nu), indexing="ij")
    a = (jj * (nu + 1) + ii).reshape(-1)
    b, c = a + 1, a + nu + 1
    d = c + 1
    t1 = np.stack([a, b, c], -1)
    t2 = np.stack([b, d, c], -1)
    vt = np.stack([t1, t2], 1).reshape(-1, 3).astype(np.int32)
    tri = np.concatenate([vt, vt, vt, np.zeros((len(vt), 1), np.int32)], 1)  # vtx, uv, normal share indices; group 0
    return vertices, normals, uvs, tri


def sky_envmap(W=2048, H=1024):
    """8-bit procedural sky: 40+60*max(0,N.y), plus a 5 degree sun disc (255) at elevation 40, azimuth 30.
    Texel (r,c) <-> direction through the reference's lookup (Geometry.h:963-977):
    theta = 1-acos(N.y)/pi = r/(H-1), phi = (atan2(-N.z,N.x)+pi)/(2pi) = c/(W-1)."""
    r = np.arange(H, dtype=np.float64)[:, None] / (H - 1)
    c = np.arange(W, dtype=np.float64)[None, :] / (W - 1)
    ny = np.cos(math.pi * (1 - r))
    s = np.sqrt(np.maximum(0, 1 - ny * ny))
    ang = 2 * math.pi * c - math.pi
    nx, nz = s * np.cos(ang), -s * np.sin(ang)
    el, az = math.radians(40), math.radians(30)
    sun = np.array([math.cos(el) * math.cos(az), math.sin(el), math.cos(el) * math.sin(az)])
    cosang = nx * sun[0] + ny * sun[1] + nz * sun[2]
    val = 40 + 60 * np.maximum(0, ny) * np.ones_like(c)
    val = np.where(cosang > math.cos(math.radians(5)), 255.0, val)
    img = np.clip(np.floor(val), 0, 255).astype(np.uint8)
    return np.repeat(img[..., None], 3, -1).copy()


def merl_table(n=60.0):
    """Synthetic MERL-format table (3 x 90 x 90 x 180 doubles): Lambert 0.2/pi + normalised Blinn-like lobe
    0.5*(n+2)/(2pi)*cos(theta_h)^n at each bin centre, divided by the channel scale so RGB come out equal
    (MERLBRDFRead.h:3-8; theta_h bins are non-linear: idx = 90*sqrt(theta_h/(pi/2)), MERLBRDFRead.cpp:130-142)."""
    ih = (np.arange(90, dtype=np.float64) + 0.5) / 90
    theta_h = ih * ih * (math.pi / 2)
    val = 0.2 / math.pi + 0.5 * (n + 2) / (2 * math.pi) * np.cos(theta_h) ** n
    tab = np.broadcast_to(val[:, None, None], (90, 90, 180))
    scales = (1.0 / 1500.0, 1.15 / 1500.0, 1.66 / 1500.0)
    return np.stack([tab / s for s in scales], 0).astype(np.float64).copy()


def wave_normal_map(size=1024):
    """Post-load normal map values: normalize(0.3 sin(40 pi u), 0.3 sin(40 pi v), 1)."""
    t = (np.arange(size, dtype=np.float64) + 0.0) / (size - 1)
    x = 0.3 * np.sin(40 * math.pi * t)[None, :] * np.ones((size, 1))
    y = 0.3 * np.sin(40 * math.pi * t)[:, None] * np.ones((1, size))
    z = np.ones((size, size))
    n = np.stack([x, y, z], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    return n.astype(np.float32)


def checker_alpha_map(size=1024):
    """Alpha 0 where (floor(16u)+floor(16v)) % 8 == 0, else 1 (post-load float RGB)."""
    t = np.arange(size, dtype=np.float64) / (size - 1)
    k = np.minimum(np.floor(16 * t), 15).astype(np.int64)
    a = ((k[None, :] + k[:, None]) % 8 != 0).astype(np.float32)
    return np.repeat(a[..., None], 3, -1).copy()


def _place_like_gui(mesh, scale=30.0, ground_y=-27.3):
    """mainApp.cpp:2402-2410: scale, then rest the bbox on the ground plane.  After TriMesh::init with
    center=true the mesh spans a unit max-extent box centred at 0; min.y follows from the vertex data."""
    v = mesh.vertices.astype(np.float32)
    # init: (x,y,z)->(-z,y,x), then (v-c)/s*scaling+offset, all float32 (TriangleMesh.cpp:742-770)
    sw = np.stack([-v[:, 2], v[:, 1], v[:, 0]], -1)
    lo, hi = sw.min(0), sw.max(0)
    s = np.float32(max(hi - lo))
    c = (lo + hi) * np.float32(0.5)
    miny = np.float32(((sw[:, 1] - c[1]) / s * np.float32(mesh.scaling) + np.float32(mesh.offset[1])).min())
    mesh.scale = float(scale)
    mesh.max_translation = np.array([0, np.float32(ground_y) - miny * np.float32(scale), 0], np.float32)
    return mesh


def phong(Kd, Ks, Ne, **extra):
    d = dict(Kd=Texture(Kd), Ks=Texture(Ks), Ne=Texture(Ne), alpha=Texture(1.0), refr=Texture(1.3),
             transp=Texture(1.0), normal=Texture((0, 0, 1)))
    d.update(extra)
    return d


def base(lib, W, H, spp, depth=5, device=0):
    rt = Raytracer(lib, device).loadScene()
    rt.W, rt.H, rt.nrays, rt.nb_bounces = W, H, spp, depth
    return rt


def config_C1(lib, W=512, H=512, spp=64, device=0):
    """Default parametric scene + two Phong spheres."""
    rt = base(lib, W, H, spp, device=device)
    s1 = Sphere((0, -17.3, 0), 10).set_material(0, **phong((.8, .3, .3), 0.0, 1.0))
    s2 = Sphere((-15, -20.3, 5), 7).set_material(0, **phong((.3, .8, .3), 0.3, 50.0))
    rt.s.addObject(s1)
    rt.s.addObject(s2)
    return rt


def config_ngan(lib, W=128, H=128, spp=8, nv=24, device=0):
    """The material presets of the reference's object menu (mainApp.cpp:1499-1597) on real objects: a torus in the Ngan fit of gold,
    spheres in the Ngan fits of copper (Ne 33200: a needle-sharp lobe) and pearl, and one in the OpenGL-table bronze."""
    rt = base(lib, W, H, spp, device=device)
    m = _place_like_gui(TriMesh(*displaced_torus(nv)))
    m.set_material(0, **phong((.5, .5, .5), (.2, .2, .2), 50.0)).set_preset("gold_ngan", 0)
    rt.s.addObject(m)
    for O, R, name in (((-20, -20.3, 8), 7, "copper_ngan"), ((19, -21.3, 10), 6, "pearl_ngan"), ((2, -23.3, 22), 4, "bronze")):
        rt.s.addObject(Sphere(O, R).set_material(0, **phong((.8, .3, .3), 0.1, 10.0)).set_preset(name, 0))
    rt.s.objects[1].envmap = sky_envmap(128, 64)
    return rt


def config_cyl(lib, W=128, H=128, spp=8, nv=16, device=0):
    """Cylinder objects (Geometry.h:731-846, the reference's yarn segments): a standing textured tube, a tilted one that is scaled and
    rotated by its object matrix, a mirror tube lying on the ground, one seen through its open end, plus a small torus that casts
    and receives their shadows."""
    rt = base(lib, W, H, spp, device=device)
    t = np.linspace(0, 1, 32, dtype=np.float64)
    stripes = np.repeat(np.repeat((0.25 + 0.7 * (np.floor(t * 8) % 2))[None, :, None], 4, 0), 3, 2).astype(np.float32)   # varies along the axis (u = dP/len)
    a = Cylinder((-14, -27.3, 6), (-14, -9, 6), 4.0).set_material(0, **phong((.9, .5, .2), 0.2, 30.0))
    a.materials[0]["Kd"] = Texture((.9, .5, .2), stripes)
    b = Cylinder((6, -25, -4), (16, -12, 4), 2.5).set_material(0, **phong((.3, .5, .9), 0.4, 80.0))
    b.scale, b.mat_rotation = 1.2, _rot(0.3, 0.5)
    b.max_translation = np.array([2, 1, -3], np.float32)
    c = Cylinder((-6, -24.8, 16), (8, -24.8, 20), 2.5, mirror=True).set_material(0, **phong((.8, .8, .8), 0.0, 1.0))
    d = Cylinder((18, -20, 26), (24, -16, 6), 3.0).set_material(0, **phong((.6, .8, .3), 0.1, 10.0))   # roughly along the view: the open end
    d.flip_normals = True
    for o in (a, b, c, d):
        rt.s.addObject(o)
    m = _place_like_gui(TriMesh(*displaced_torus(nv)), scale=14.0)
    m.max_translation = m.max_translation + np.array([0, 0, 2], np.float32)
    m.set_material(0, **phong((.5, .5, .5), (.2, .2, .2), 50.0))
    rt.s.addObject(m)
    return rt


def torus_points(nv):
    """The vertices of displaced_torus(nv) as an oriented point set in the state PointSet::init leaves: positions centred and
    normalised to a unit extent, analytic normals, one radius per point from the local vertex spacing, colours from the position."""
    v, n, uv, _ = displaced_torus(nv)
    v, n = v.astype(np.float64), n.astype(np.float64)
    lo, hi = v.min(0), v.max(0)
    p = (v - (lo + hi) / 2) / (hi - lo).max()
    grid = p.reshape(nv + 1, 2 * nv + 1, 3)
    du = np.linalg.norm(np.diff(grid, axis=1, append=grid[:, -2:-1]), axis=-1)
    dv = np.linalg.norm(np.diff(grid, axis=0, append=grid[-2:-1]), axis=-1)
    rad = 0.75 * np.maximum(du, dv)
    col = np.stack([0.25 + 0.7 * uv[:, 0], 0.3 + 0.6 * uv[:, 1], 0.9 - 0.6 * uv[:, 0]], -1)
    keep = np.zeros((nv + 1, 2 * nv + 1), bool)
    keep[:-1, :-1] = True                         # the seam rows / columns repeat the first ones: coincident discs would tie
    keep = keep.reshape(-1)
    return p[keep].astype(np.float32), n[keep].astype(np.float32), rad.reshape(-1)[keep].astype(np.float32), col[keep].astype(np.float32)


def config_points(lib, W=128, H=128, spp=8, nv=40, device=0, display_edges=False):
    """A PointSet (PointSet.cpp: discs in the object's own space, its colour per point, normals turned to the viewer) placed like the
    GUI places a mesh, a second, mirror-flagged coarse one, next to a sphere that casts and receives shadows with them."""
    rt = base(lib, W, H, spp, device=device)
    ps = PointSet(*torus_points(nv), display_edges=display_edges)
    ps.set_material(0, **phong((.5, .5, .5), (.15, .15, .15), 40.0))
    ps.scale = 30.0
    miny = float(ps.points[:, 1].min())
    ps.max_translation = np.array([0, np.float32(-27.3) - np.float32(miny) * np.float32(30.0), 0], np.float32)
    ps.mat_rotation = _rot(0.0, 0.6)
    rt.s.addObject(ps)
    small = PointSet(*torus_points(max(6, nv // 4))[:3], mirror=False, normal_swapped=True)
    small.set_material(0, **phong((.5, .5, .5), 0.0, 1.0))
    small.scale = 10.0
    small.max_translation = np.array([-22, -20, 10], np.float32)
    rt.s.addObject(small)
    rt.s.addObject(Sphere((17, -21.3, 14), 6).set_material(0, **phong((.3, .8, .3), 0.3, 50.0)))
    return rt


def weave_segments(n_warp=6, n_weft=6, seg=20, radius=0.028):
    """A plain-weave patch in the unit square of the xz plane as yarn segments (what Yarns::cyls holds): `n_warp` curves along x and
    `n_weft` along z, each passing over and under the ones it crosses; radii alternate per yarn."""
    A, B, R = [], [], []
    amp = 1.1 * radius
    for fam, n_own, n_cross in ((0, n_warp, n_weft), (1, n_weft, n_warp)):
        for k in range(n_own):
            s = np.linspace(-0.5, 0.5, seg + 1)
            off = (k + 0.5) / n_own - 0.5
            phase = np.pi * (k + fam)
            y = amp * np.cos(np.pi * n_cross * (s + 0.5) + phase)
            pts = np.stack([s, y, np.full_like(s, off)], -1) if fam == 0 else np.stack([np.full_like(s, off), y, s], -1)
            A.append(pts[:-1]); B.append(pts[1:])
            R.append(np.full(seg, radius * (1.0 if k % 2 == 0 else 0.8)))
    return np.concatenate(A).astype(np.float32), np.concatenate(B).astype(np.float32), np.concatenate(R).astype(np.float32)


def config_yarns(lib, W=128, H=128, spp=8, seg=20, device=0):
    """Yarns objects (TriangleMesh.h:265-312): a tilted, scaled weave of yarn segments over the ground, a second, mirror-flagged
    coarse one, next to a sphere that casts and receives shadows with them."""
    rt = base(lib, W, H, spp, device=device)
    # (the sphere first: an object tested after the yarns would reset the picking query's triangle id whenever the ray meets it at
    #  all, nearer or not: Sphere::intersection writes triangle_id = -1, Geometry.h:990)
    rt.s.addObject(Sphere((-17, -21.3, 14), 6).set_material(0, **phong((.3, .8, .3), 0.3, 50.0)))
    y = Yarns(*weave_segments(6, 6, seg))
    y.scale = 34.0
    y.mat_rotation = _rot(0.9, 0.5)
    y.max_translation = np.array([-3, -12, 4], np.float32)
    rt.s.addObject(y)
    small = Yarns(*weave_segments(3, 2, max(4, seg // 3), radius=0.06), mirror=True)
    small.scale = 14.0
    small.mat_rotation = _rot(0.2, 1.1)
    small.max_translation = np.array([20, -19, 12], np.float32)
    rt.s.addObject(small)
    return rt


def config_C2(lib, W=1024, H=1024, spp=256, nv=500, env=(2048, 1024), device=0):
    """1M-triangle Phong torus + 8-bit sky envmap."""
    rt = base(lib, W, H, spp, device=device)
    m = _place_like_gui(TriMesh(*displaced_torus(nv)))
    m.set_material(0, **phong((.5, .5, .5), (.2, .2, .2), 50.0))
    rt.s.addObject(m)
    rt.s.objects[1].envmap = sky_envmap(*env)
    return rt


def config_C3(lib, W=1920, H=1080, spp=512, nv=791, tex=1024, device=0):
    """2.5M-triangle fully transparent torus (refr 1.5) + normal and alpha maps."""
    rt = base(lib, W, H, spp, device=device)
    m = _place_like_gui(TriMesh(*displaced_torus(nv)))
    m.set_material(0, **phong((.5, .5, .5), (.2, .2, .2), 50.0, transp=Texture(0.0), refr=Texture(1.5),
                              normal=Texture((0, 0, 1), wave_normal_map(tex)), alpha=Texture(1.0, checker_alpha_map(tex))))
    rt.s.addObject(m)
    return rt


def config_C4(lib, W=1024, H=1024, spp=1024, nv=255, device=0):
    """260k-triangle MERL torus with depth of field."""
    rt = base(lib, W, H, spp, device=device)
    rt.cam.aperture, rt.cam.focus_distance = np.float32(1.0), np.float32(50)
    m = _place_like_gui(TriMesh(*displaced_torus(nv)))
    m.set_material(0, **phong((.5, .5, .5), 0.0, 1.0))
    m.brdf = ("merl", merl_table())
    rt.s.addObject(m)
    return rt


def config_C5(lib, W=3840, H=2160, spp=1024, nv=866, device=0, tori=range(8)):
    """8 tori (24M triangles at nv=866) on a 4x2 grid, scale 15, Phong, no envmap.  `tori`: which of the eight to place."""
    rt = base(lib, W, H, spp, device=device)
    geo = displaced_torus(nv)
    for k in tori:
        m = _place_like_gui(TriMesh(*geo), scale=15.0)
        gx, gz = k % 4, k // 4
        m.max_translation = m.max_translation + np.array([(gx - 1.5) * 17.5, 0, (gz - 0.5) * 17.5 - 10], np.float32)
        m.set_material(0, **phong((.5, .5, .5), (.2, .2, .2), 50.0))
        rt.s.addObject(m)
    return rt


def photo_background(W=96, H=64):
    """Closed-form stand-in for a background photograph, in Scene::background scale (pow(v/255, 2.2) * 196964.699,
    Geometry.h:1355-1363): a two-colour gradient with a bright band."""
    y = np.arange(H, dtype=np.float64)[:, None] / (H - 1)
    x = np.arange(W, dtype=np.float64)[None, :] / (W - 1)
    v = np.stack([60 + 150 * x * np.ones_like(y), 90 + 100 * y * np.ones_like(x), 200 - 120 * x * y], -1)
    v = np.where((np.abs(y - 0.5) < 0.06)[..., None], 240.0, v)
    return (np.power(np.floor(v) / 255., 2.2) * 196964.699).astype(np.float32)


def config_ghost(lib, W=128, H=128, spp=16, nv=24, device=0, background=True, ghost_plane=True, ghost_mesh=False):
    """Compositing set-up of the reference (Raytracer.cpp:260-268, 522-537, 614-621): the ground plane is a ghost that only
    receives shadows, a background photograph shows through it and behind the scene; a Phong torus and a sphere cast the shadows."""
    rt = base(lib, W, H, spp, device=device)
    rt.s.objects[2].ghost = ghost_plane
    rt.s.objects[2].set_material(0, **phong((.7, .7, .7), 0.0, 1.0))
    m = _place_like_gui(TriMesh(*displaced_torus(nv)))
    m.set_material(0, **phong((.5, .5, .5), (.2, .2, .2), 50.0))
    m.ghost = ghost_mesh
    rt.s.addObject(m)
    rt.s.addObject(Sphere((-18, -20.3, 8), 7).set_material(0, **phong((.3, .8, .3), 0.3, 50.0)))
    if background:
        rt.s.background_values = photo_background()
    return rt


def config_fog(lib, W=128, H=128, spp=16, nv=24, device=0, fog_type=0, phase=0, mesh=True):
    """Participating medium (Raytracer::fogContribution, Raytracer.cpp:40-192) over the C1/C2 content: uniform (type 0) or
    height-exponential (type 1) fog with an isotropic / Schlick / Rayleigh phase function."""
    rt = base(lib, W, H, spp, device=device)
    rt.s.addObject(Sphere((0, -17.3, 0), 10).set_material(0, **phong((.8, .3, .3), 0.0, 1.0)))
    rt.s.addObject(Sphere((-15, -20.3, 5), 7, mirror=True))
    if mesh:
        m = _place_like_gui(TriMesh(*displaced_torus(nv)), scale=20.0)
        m.max_translation = m.max_translation + np.array([16, 0, 4], np.float32)
        m.set_material(0, **phong((.5, .5, .5), (.2, .2, .2), 50.0))
        rt.s.addObject(m)
    s = rt.s
    s.fog_density, s.fog_absorption, s.fog_density_decay, s.fog_absorption_decay = 0.3, 0.3, 0.05, 0.05
    s.fog_type, s.fog_phase_type, s.phase_aniso = fog_type, phase, 0.4
    return rt


def config_exotic_modes(lib, W=128, H=128, spp=16, device=0, mode="fog"):
    """The branching modes of getColor over the primitives of row f4: yarns, a point set and a cylinder in a height-exponential fog
    (mode "fog"), or as / next to ghost objects over a ghost ground plane with a background photograph (mode "ghost": the yarns
    and the cylinder are ghosts, so shadow rays pass through them and every hit on them spawns the straight-through ray); mode
    "plain" is the same content on the linear path, mode "merl" gives the three of them the measured BRDF of C4."""
    ghost = mode == "ghost"
    rt = base(lib, W, H, spp, device=device)
    rt.s.addObject(Sphere((-17, -21.3, 14), 6).set_material(0, **phong((.3, .8, .3), 0.3, 50.0)))
    y = Yarns(*weave_segments(5, 5, 10))
    y.scale, y.mat_rotation, y.max_translation = 30.0, _rot(0.9, 0.5), np.array([-3, -13, 4], np.float32)
    y.ghost = ghost
    rt.s.addObject(y)
    ps = PointSet(*torus_points(14))
    ps.set_material(0, **phong((.5, .5, .5), (.15, .15, .15), 40.0))
    ps.scale, ps.mat_rotation = 14.0, _rot(0.3, 0.9)
    ps.max_translation = np.array([19, -20, 10], np.float32)
    rt.s.addObject(ps)
    cy = Cylinder((2, -27.3, 22), (6, -15, 20), 2.0).set_material(0, **phong((.9, .5, .2), 0.2, 30.0))
    cy.ghost = ghost
    rt.s.addObject(cy)
    if ghost:
        rt.s.objects[2].ghost = True
        rt.s.objects[2].set_material(0, **phong((.7, .7, .7), 0.0, 1.0))
        rt.s.background_values = photo_background()
    elif mode == "merl":      # the BRDF is the OBJECT's: measured (IsoMERLBRDF) yarns, discs and cylinder, with depth of field
        rt.cam.aperture, rt.cam.focus_distance = np.float32(0.6), np.float32(50)
        table = merl_table()
        for o in (y, ps, cy):
            o.brdf = ("merl", table)
    elif mode == "fog":
        s = rt.s
        s.fog_density, s.fog_absorption, s.fog_density_decay, s.fog_absorption_decay = 0.3, 0.3, 0.05, 0.05
        s.fog_type, s.fog_phase_type, s.phase_aniso = 1, 1, 0.4
    return rt


def config_sss(lib, W=128, H=128, spp=16, nv=24, device=0, mixed=True):
    """Subsurface scattering (Raytracer.cpp:318-406) on a torus: Ksub drives the below-surface random walk step; a second,
    ordinary Phong torus and a mirror sphere share the scene when `mixed`."""
    rt = base(lib, W, H, spp, device=device)
    m = _place_like_gui(TriMesh(*displaced_torus(nv)))
    m.set_material(0, **phong((.5, .45, .4), (.1, .1, .1), 30.0, Ksub=Texture((.8, .5, .4))))
    rt.s.addObject(m)
    if mixed:
        m2 = _place_like_gui(TriMesh(*displaced_torus(max(8, nv // 2))), scale=12.0)
        m2.max_translation = m2.max_translation + np.array([-20, 0, 10], np.float32)
        m2.set_material(0, **phong((.3, .6, .8), (.2, .2, .2), 50.0))
        rt.s.addObject(m2)
        rt.s.addObject(Sphere((18, -21.3, 12), 6, mirror=True))
    return rt


def _rot(ax, ay):
    cx, sx, cy, sy = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    return (ry @ rx).astype(np.float32)


def config_anim(lib, W=128, H=128, spp=8, nv=20, frame=0, device=0):
    """Key-framed placement (Object::get_translation / get_rotation / get_scale + Slerp, Geometry.h:258-312): a torus that moves,
    turns and shrinks between frames 2 and 8, a sphere on a three-key translation track, and a light whose radius is keyed."""
    rt = base(lib, W, H, spp, device=device)
    m = _place_like_gui(TriMesh(*displaced_torus(nv)))
    m.set_material(0, **phong((.5, .5, .5), (.2, .2, .2), 50.0))
    m.add_keyframe(2)
    m.scale, m.mat_rotation = 20.0, _rot(0.7, 0.9)
    m.max_translation = m.max_translation + np.array([8, 4, -6], np.float32)
    m.add_keyframe(8)
    # (the sphere goes in before the mesh: Sphere::intersection resets `triangle_id` to -1 whenever it is hit, even behind a closer
    #  mesh, so the reference's picking query loses the triangle id of mesh pixels in front of a LATER sphere, Geometry.h:990)
    sp = Sphere((-15, -20.3, 5), 7).set_material(0, **phong((.3, .8, .3), 0.3, 50.0))
    for fr, t in ((0, (0, 0, 0)), (4, (6, 3, 0)), (9, (6, 12, -8))):
        sp.translation_keyframes[float(fr)] = np.array(t, np.float32)
        sp.scale_keyframes[float(fr)] = 1.0
        sp.rotation_keyframes[float(fr)] = np.eye(3, dtype=np.float32)
    rt.s.addObject(sp)
    rt.s.addObject(m)
    light = rt.s.objects[0]
    for fr, sc in ((1, 1.0), (7, 0.5)):
        light.scale_keyframes[float(fr)] = sc
        light.translation_keyframes[float(fr)] = np.array((0, 0, -2.0 * fr), np.float32)
        light.rotation_keyframes[float(fr)] = np.eye(3, dtype=np.float32)
    rt.s.current_frame = frame
    return rt


CONFIGS = {"C1": config_C1, "C2": config_C2, "C3": config_C3, "C4": config_C4, "C5": config_C5}
