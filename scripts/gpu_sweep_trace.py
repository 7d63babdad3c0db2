"""Sweep of the traversal knobs (refill threshold, triangle-phase fraction) with the current defaults otherwise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, _abi
G = ptb.load()
for spec in (sys.argv[1:] or ["C2:128", "C3:64"]):
    wl, spp = spec.split(":")
    rt = scenes.CONFIGS[wl](G); rt.nrays = int(spp); rt.commit()
    for refill, den, pct in [tuple(int(x) for x in a.split(',')) for a in os.environ.get('PTB_SWEEP', '24,4,25 24,4,25 24,4,0 24,4,15 24,4,35 22,4,25 26,4,25 24,3,25 24,6,25').split()]:
        rt.set_option(_abi.OPT_REFILL_BELOW, refill); rt.set_option(_abi.OPT_TRI_FRACTION, den); rt.set_option(_abi.OPT_TRI_MIN_PCT, pct)
        best = 1e30
        for rep in range(3):
            rt.render_image_nopreviz(want_image=False); best = min(best, rt.stats["ms_device"])
        print(f"{wl} spp={spp} refill<{refill} tri_den={den} tri_min_pct={pct}: {best:8.2f} ms", flush=True)
    rt.close()
