"""Sweep the wavefront pool size on the GPU box (does an L2-resident pool pay?)."""
import itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, _abi
G = ptb.load()
for wl, spp in (("C2", 64), ("C3", 32), ("C4", 64)):
    rt = scenes.CONFIGS[wl](G); rt.nrays = spp; rt.commit()
    rt.set_option(_abi.OPT_TIME_KERNELS, 1)
    rt.render_image_nopreviz(want_image=False)
    for logp in (19, 20, 21, 22, 23, 24, 25):
        rt.set_option(_abi.OPT_POOL_PATHS, 1 << logp)
        best = None
        for _ in range(2):
            rt.render_image_nopreviz(want_image=False)
            kt = rt.kernel_times(); s = rt.stats
            row = dict(wl=wl, pool_log2=logp, ms=round(s["ms_device"], 2), extend=round(kt["extend"]["ms"], 2), shade=round(kt["shade"]["ms"], 2), shadow=round(kt["shadow"]["ms"], 2),
                       raygen=round(kt["raygen"]["ms"], 2), splat=round(kt["splat"]["ms"], 2), launches=s["kernel_launches"],
                       msamples=round(s["samples"] / s["ms_device"] / 1e3, 1), mean=round(float(rt.imagedouble.mean()), 2))
            if best is None or row["ms"] < best["ms"]: best = row
        print(json.dumps(best), flush=True)
    rt.close()
