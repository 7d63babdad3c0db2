"""GPU diagnostic: the YARN golden scene on the device against the oracle port: where do the images differ?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, _abi
from oracles import port_lib
from parity_cases import rel_err

G, P = ptb.load(), port_lib()
tag = sys.argv[1] if len(sys.argv) > 1 else "intree"
def run(mk, label, nb=None):
    a, b = mk(P), mk(G)
    if nb is not None: a.nb_bounces = b.nb_bounces = nb
    a.commit(); b.commit()
    oa, ob = a.primary_ids(), b.primary_ids()
    ia, ib = a.render_image_nopreviz().copy(), b.render_image_nopreviz().copy()
    e = rel_err(ia, ib)
    bad = e > 1e-3
    by = {int(k): (int((bad & (oa[0] == k)).sum()), int((oa[0] == k).sum())) for k in np.unique(oa[0])}
    print(f"[{tag}] {label}: ids agree {((oa[0]==ob[0])&(oa[1]==ob[1])).mean():.5f}  bad {bad.mean():.5f}  mean {ia.mean():.3f} / {ib.mean():.3f}  by primary object (bad, pixels): {by}"
          f"  rays closest {a.stats['rays_closest']} / {b.stats['rays_closest']} shadow {a.stats['rays_shadow']} / {b.stats['rays_shadow']}", flush=True)
    a.close(); b.close()

for nb in (1, 2, 5):
    run(lambda L: scenes.config_yarns(L, 48, 48, 2, seg=12), f"YARN nb={nb}", nb)
def no_small(L):
    rt = scenes.config_yarns(L, 48, 48, 2, seg=12); rt.s.objects.pop(5); return rt
def no_big(L):
    rt = scenes.config_yarns(L, 48, 48, 2, seg=12); rt.s.objects.pop(4); return rt
run(no_small, "YARN without the mirror yarns")
run(no_big, "YARN without the big weave")
run(lambda L: scenes.config_yarns(L, 128, 128, 4, seg=20), "YARN 128x128x4")
run(lambda L: scenes.config_points(L, 48, 48, 2, nv=24), "PTS (passes)")
