"""The miniature scenes and KAT inputs shared by tests/golden/make_golden.py and the tests."""
import numpy as np

from pathtracer_b200 import _abi, scenes

SCENES = {
    "C1": lambda L, **kw: scenes.config_C1(L, 48, 48, 2, **kw),
    "C2": lambda L, **kw: scenes.config_C2(L, 48, 48, 2, nv=24, env=(128, 64), **kw),
    "C3": lambda L, **kw: scenes.config_C3(L, 48, 48, 2, nv=24, tex=64, **kw),
    "C4": lambda L, **kw: scenes.config_C4(L, 48, 48, 2, nv=24, **kw),
    "C5": lambda L, **kw: scenes.config_C5(L, 64, 40, 1, nv=12, **kw),
    "NGAN": lambda L, **kw: scenes.config_ngan(L, 48, 48, 2, **kw),
    "CYL": lambda L, **kw: scenes.config_cyl(L, 48, 48, 2, **kw),
    "PTS": lambda L, **kw: scenes.config_points(L, 48, 48, 2, nv=24, **kw),                        # PointSet discs (PointSet.cpp)
    "PTS_EDGES": lambda L, **kw: scenes.config_points(L, 48, 48, 2, nv=16, display_edges=True, **kw),
    "YARN": lambda L, **kw: scenes.config_yarns(L, 48, 48, 2, seg=12, **kw),                       # Yarns (TriangleMesh.h:265-312, TriangleMesh.cpp:1519-1737)
    "MERL_EXOTIC": lambda L, **kw: scenes.config_exotic_modes(L, 48, 48, 2, mode="merl", **kw),     # IsoMERLBRDF on yarns, discs and a cylinder, thin lens
}

# getColor's branching modes: background photograph, ghost objects, participating medium
BRANCH_SCENES = {
    "BG": lambda L, **kw: scenes.config_ghost(L, 48, 48, 2, ghost_plane=False, **kw),
    "GHOST": lambda L, **kw: scenes.config_ghost(L, 48, 48, 2, **kw),
    "GHOSTMESH": lambda L, **kw: scenes.config_ghost(L, 48, 48, 2, ghost_mesh=True, **kw),
    "GHOSTNOBG": lambda L, **kw: scenes.config_ghost(L, 48, 48, 2, background=False, **kw),
    "FOG_U0": lambda L, **kw: scenes.config_fog(L, 48, 48, 2, fog_type=0, phase=0, **kw),
    "FOG_U1": lambda L, **kw: scenes.config_fog(L, 48, 48, 2, fog_type=0, phase=1, mesh=False, **kw),
    "FOG_E1": lambda L, **kw: scenes.config_fog(L, 48, 48, 2, fog_type=1, phase=1, **kw),
    "FOG_E2": lambda L, **kw: scenes.config_fog(L, 48, 48, 2, fog_type=1, phase=2, **kw),
    "FOG_EXOTIC": lambda L, **kw: scenes.config_exotic_modes(L, 48, 48, 2, mode="fog", **kw),        # yarns, discs and a cylinder in the medium
    "GHOST_EXOTIC": lambda L, **kw: scenes.config_exotic_modes(L, 48, 48, 2, mode="ghost", **kw),    # ghost yarns / cylinder over a ghost ground + photograph
}
# key-framed placement at several frames: before the first key, between keys (Slerp), on a key, after the last
ANIM_SCENES = {f"ANIM_F{fr}": (lambda L, fr=fr, **kw: scenes.config_anim(L, 48, 48, 2, frame=fr, **kw)) for fr in (0, 3, 5, 8, 12)}

# the subsurface branch (Raytracer.cpp:318-406): pinned port == reference bit for bit; the CUDA path can only be compared
# statistically (the reservoir over a probe ray's hits draws in BVH traversal order, parity_cases.case_sss_converged)
STAT_SCENES = {
    "SSS": lambda L, **kw: scenes.config_sss(L, 48, 48, 2, **kw),
    "SSS_ALONE": lambda L, **kw: scenes.config_sss(L, 48, 48, 2, mixed=False, **kw),
}

def _unit(rng, n):
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def KAT_INPUTS(n=256):
    rng = np.random.default_rng(20261017)
    N, wi, wo = _unit(rng, n), _unit(rng, n), _unit(rng, n)
    # MERL: directions in the upper hemisphere of N so that most lookups land inside the table
    wi_up = wi * np.sign((wi * N).sum(1, keepdims=True))
    wo_up = wo * np.sign((wo * N).sum(1, keepdims=True))
    return {
        _abi.KAT_PCG32: (np.stack([rng.integers(0, 2 ** 40, n), rng.integers(0, 2 ** 33, n)], 1).astype(np.float64), {}),
        _abi.KAT_LATTICE: (np.arange(n, dtype=np.float64)[:, None], {}),
        _abi.KAT_CAMERA: (np.concatenate([rng.integers(0, 512, (n, 2)), rng.uniform(-.5, .5, (n, 2)), rng.uniform(-.05, .05, (n, 2))], 1), dict(W=512, H=512)),
        _abi.KAT_RANDOM_COS: (np.concatenate([N, rng.uniform(0, 1, (n, 2))], 1), {}),
        _abi.KAT_RANDOM_PHONG: (np.concatenate([N, rng.uniform(1, 100, (n, 1)), rng.uniform(0, 1, (n, 2))], 1), {}),
        _abi.KAT_PHONG_EVAL: (np.concatenate([rng.uniform(0, 1, (n, 6)), rng.uniform(1, 100, (n, 3)), wi, wo, N], 1), {}),
        _abi.KAT_MERL_EVAL: (np.concatenate([wi_up, wo_up, N], 1), {}),
        _abi.KAT_FAST_EXP: (rng.uniform(-8, 0, (n, 1)), {}),
        _abi.KAT_FAST_NORMALIZE: (rng.normal(size=(n, 3)) * 10, {}),
        _abi.KAT_RANDOM_PER_PIXEL: (rng.integers(0, 96 * 96, (64, 1)).astype(np.float64), dict(W=96, H=96)),
        _abi.KAT_FILTER_RATIO: (np.array([[0, 0, .5], [5, 0, .5], [95, 95, .5], [60, 60, .5], [0, 95, .5], [0, 0, 1.0], [50, 95, 1.0], [40, 40, 1.0]]), dict(W=96, H=96)),
    }
