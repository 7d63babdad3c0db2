"""Per-bounce launch times of one pass (PTB_DEBUG_BOUNCES) on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PTB_DEBUG_BOUNCES"] = "1"
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, _abi
G = ptb.load()
for wl, spp in (("C2", 32), ("C3", 16), ("C4", 32)):
    rt = scenes.CONFIGS[wl](G); rt.nrays = spp; rt.commit()
    rt.set_option(_abi.OPT_TIME_KERNELS, 1)
    os.environ.pop("PTB_DEBUG_BOUNCES", None)
    rt.render_image_nopreviz(want_image=False)
    os.environ["PTB_DEBUG_BOUNCES"] = "1"
    print("==", wl, "paths per pass", rt.W * rt.H * min(spp, (1 << 25) // (rt.W * rt.H)), flush=True)
    sys.stderr.flush()
    rt.render_image_nopreviz(want_image=False)
    rt.close()
