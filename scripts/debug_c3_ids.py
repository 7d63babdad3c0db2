import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import pathtracer_b200
from pathtracer_b200 import scenes
from oracles import port_lib
a = scenes.config_C3(port_lib(), spp=1).commit(); b = scenes.config_C3(pathtracer_b200.load(), spp=1).commit()
oa, ta, da = a.primary_ids(); ob, tb, db = b.primary_ids()
bad = (oa != ob) | (ta != tb)
print("mismatch", int(bad.sum()), "of", bad.size, "obj differs", int((oa != ob).sum()))
rel = np.abs(da - db) / np.maximum(np.abs(da), 1e-6)
same_t = bad & (rel < 1e-5)
print("mismatch with equal t (<1e-5 rel):", int(same_t.sum()), " with t differing:", int((bad & ~same_t).sum()))
idx = np.argwhere(bad & ~same_t)[:12]
for i, j in idx: print(i, j, "port", oa[i, j], ta[i, j], da[i, j], "gpu", ob[i, j], tb[i, j], db[i, j])
d = np.abs(ta - tb)[same_t]
print("id distance histogram (equal-t mismatches):", np.unique(np.minimum(d, 5000), return_counts=True))
