import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import pathtracer_b200
from pathtracer_b200 import scenes
from oracles import port_lib
out = {}
for name, L in (("gpu", pathtracer_b200.load()), ("port", port_lib())):
    rt = scenes.config_C4(L, spp=4).commit()
    img = rt.render_image_nopreviz().copy()
    neg = np.argwhere(img.min(-1) < 0)
    out[name] = img
    print(name, "min", float(img.min()), "neg px", len(neg), "mean", float(img.mean()), [(int(i), int(j), float(img[i, j, 0])) for i, j in neg[:8]], flush=True)
a, b = out["port"], out["gpu"]
rel = np.abs(a - b).max(-1) / np.maximum(np.abs(a).max(-1), 1e-3 * float(a.mean()))
print("pixels off by >1e-3:", float(np.mean(rel > 1e-3)))
neg = np.argwhere(b.min(-1) < 0)
for i, j in neg[:8]: print(i, j, "gpu", b[i, j], "port", a[i, j])
