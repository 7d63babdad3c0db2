"""One process, N GPUs (ptb_group_*): the reference-shaped single call rendered by all the GPUs of the box.
usage: python scripts/gpu_group_bench.py [C2:256 C5:64]  ->  Msamples/s with host outputs (pinned), commit time, per workload and N"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes

G = ptb.load()
n_all = torch.cuda.device_count()
for spec in ([a for a in sys.argv[1:] if ":" in a] or ["C2:256", "C5:64"]):
    wl, spp = spec.split(":")
    for n in sorted({1, n_all}):
        rt = scenes.CONFIGS[wl](G)
        rt.nrays = int(spp)
        rt.devices = list(range(n)) if n > 1 else None
        rt.reuse_buffers = True
        t0 = time.time(); rt.commit(); commit_s = time.time() - t0
        for _ in range(2):
            rt.render_image_nopreviz()
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter(); rt.render_image_nopreviz(); best = min(best, time.perf_counter() - t0)
        print(f"[group] {wl} spp={spp} gpus={n}: commit {commit_s:6.2f} s  render {best * 1e3:9.2f} ms  {rt.W * rt.H * rt.nrays / best / 1e6:9.1f} Msamples/s (host outputs)  "
              f"mean={float(rt.imagedouble.mean()):.3f}", flush=True)
        rt.close()
