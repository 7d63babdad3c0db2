"""Sweep the traversal-kernel knobs on the GPU box (refill threshold, triangle fraction, grid size)."""
import itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, _abi
G = ptb.load()
for wl, spp in (("C3", 16), ("C2", 32)):
    rt = scenes.CONFIGS[wl](G); rt.nrays = spp; rt.commit()
    rt.set_option(_abi.OPT_TIME_KERNELS, 1)
    rt.render_image_nopreviz(want_image=False)
    for refill, den in itertools.product((8, 16, 20, 24, 28, 32), (1, 2, 4, 8, 64)):
        rt.set_option(_abi.OPT_REFILL_BELOW, refill); rt.set_option(_abi.OPT_TRI_FRACTION, den)
        best = None
        for _ in range(2):
            rt.render_image_nopreviz(want_image=False)
            kt = rt.kernel_times(); s = rt.stats
            row = dict(wl=wl, refill=refill, den=den, ms=round(s["ms_device"], 2), extend=round(kt["extend"]["ms"], 2), shade=round(kt["shade"]["ms"], 2), shadow=round(kt["shadow"]["ms"], 2),
                       msamples=round(s["samples"] / s["ms_device"] / 1e3, 1))
            if best is None or row["ms"] < best["ms"]: best = row
        print(json.dumps(best), flush=True)
    rt.close()
