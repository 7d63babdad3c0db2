"""Sweep the traversal-kernel knobs on the GPU box; every configuration is checked against the default's image mean."""
import itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, _abi
G = ptb.load()
grid = list(itertools.product((16, 20, 24), (4,), (0, 20, 30, 40, 50, 60, 75)))
for wl, spp in (("C3", 16), ("C2", 32), ("C4", 16)):
    rt = scenes.CONFIGS[wl](G); rt.nrays = spp; rt.commit()
    rt.set_option(_abi.OPT_TIME_KERNELS, 1); rt.set_option(_abi.OPT_COUNT_TRAVERSAL, int(os.environ.get("PTB_COUNT", "0")))
    rt.render_image_nopreviz(want_image=False)
    for refill, den, pct in grid:
        rt.set_option(_abi.OPT_REFILL_BELOW, refill); rt.set_option(_abi.OPT_TRI_FRACTION, den); rt.set_option(_abi.OPT_TRI_MIN_PCT, pct)
        best = None
        for _ in range(2):
            rt.render_image_nopreviz(want_image=False)
            kt = rt.kernel_times(); s = rt.stats
            rays = s["rays_closest"] + s["rays_shadow"]
            row = dict(wl=wl, refill=refill, den=den, pct=pct, ms=round(s["ms_device"], 2), extend=round(kt["extend"]["ms"], 2), shade=round(kt["shade"]["ms"], 2), shadow=round(kt["shadow"]["ms"], 2),
                       msamples=round(s["samples"] / s["ms_device"] / 1e3, 1), mean=round(float(rt.imagedouble.mean()), 2), nodes=round(s["node_visits"] / rays, 2), tris=round(s["tri_tests"] / rays, 2))
            if best is None or row["ms"] < best["ms"]: best = row
        print(json.dumps(best), flush=True)
    rt.close()
