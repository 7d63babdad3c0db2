#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ from oracle/_ref — the REFERENCE's own sources
compiled headless by oracle/build_ref.py (only possible where /root/reference is mounted).

The fixtures pin oracle/port (and through it the CUDA path) to outputs of the reference itself:
  kat.npz        function-level known answers: inputs + reference outputs for every PTB_KAT_* block
  scene_<n>.npz  for five miniature versions of the BASELINE.json configurations: primary-hit object /
                 triangle ids and t (the picking query), the linear image (imagedouble), sample_count,
                 the 8-bit image and the ray counters of a single-thread render
Run:  python tests/golden/make_golden.py        (rewrites the .npz files; they are committed)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from golden_scenes import KAT_INPUTS, SCENES  # noqa: E402
from oracles import ref_lib  # noqa: E402

from pathtracer_b200 import _abi, scenes  # noqa: E402


def main():
    R = ref_lib()
    assert R is not None, "oracle/_ref is not built (needs /root/reference)"
    rt = scenes.config_C4(R, 32, 32, 1, nv=10).commit()   # any committed scene with a MERL table
    out = {}
    for which, (inp, kw) in KAT_INPUTS().items():
        out[f"in_{which}"] = inp
        out[f"out_{which}"] = rt.kat(which, inp, **kw)
    np.savez_compressed(os.path.join(HERE, "kat.npz"), **out)
    for name, mk in SCENES.items():
        rt = mk(R).commit()
        rt.set_option(_abi.ORC_OPT_THREADS, 1)
        obj, tri, t = rt.primary_ids()
        img = rt.render_image_nopreviz().copy()
        np.savez_compressed(os.path.join(HERE, f"scene_{name}.npz"), obj=obj.astype(np.int16), tri=tri, t=t, imagedouble=img,
                            sample_count=rt.sample_count, image=rt.image, rays=np.array([rt.stats["rays_closest"], rt.stats["rays_shadow"]], np.int64))
        print(name, img.shape, float(img.mean()), rt.stats["rays_closest"], rt.stats["rays_shadow"])


if __name__ == "__main__":
    main()
