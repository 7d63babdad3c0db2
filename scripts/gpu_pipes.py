"""Sweep of the pass pipelines (PTB_OPT_PIPES) x pool size x persistent-grid size on the GPU box: device ms of one render
(CUDA events on the context's stream, pipelines joined before the stop event), best of `reps`.
usage: python scripts/gpu_pipes.py [C2:256 C3:128 C4:256] > gpurun_out/pipes.log"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, _abi

G = ptb.load()
specs = [a for a in sys.argv[1:] if ":" in a] or ["C2:256", "C3:128", "C4:256"]
SETTINGS = [  # (pipes, log2 pool per pipe, trace blocks per SM or 0 = default)
    (1, 25, 0), (1, 25, 0),
    (2, 24, 0), (2, 25, 0), (2, 24, 5), (2, 24, 6), (2, 24, 7), (2, 25, 6),
    (3, 24, 0), (3, 24, 4), (3, 24, 5), (4, 23, 0), (4, 24, 4), (4, 23, 3),
]
for spec in specs:
    wl, spp = spec.split(":")
    rt = scenes.CONFIGS[wl](G); rt.nrays = int(spp); rt.commit()
    ref = None
    for pipes, lp, tb in SETTINGS:
        rt.set_option(_abi.OPT_PIPES, pipes)
        rt.set_option(_abi.OPT_POOL_PATHS, 1 << lp)
        rt.set_option(_abi.OPT_TRACE_BLOCKS, 148 * (tb if tb else 9))
        best = 1e30
        for rep in range(3):
            img = rt.render_image_nopreviz(want_image=False)
            best = min(best, rt.stats["ms_device"])
        samples = rt.W * rt.H * rt.nrays
        mean = float(img.mean())
        if ref is None:
            ref = mean
        print(f"{wl} spp={spp} pipes={pipes} pool=2^{lp} trace_blocks/SM={tb or 9}: {best:9.2f} ms  {samples / best / 1e3:8.1f} Msamples/s  "
              f"launches={rt.stats['kernel_launches']} mean={mean:.4f} (d {abs(mean / ref - 1):.1e})", flush=True)
    rt.close()
