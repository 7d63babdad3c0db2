"""GPU diagnostic: the yarn cloth (tests/parity_cases.yarn_cloth_scene) against the oracle port at several radii: primary ids and the
equal-seed image bound.  Thin, far tubes sit in the regime where Cylinder::intersection's own float arithmetic is noise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pathtracer_b200 as ptb
from oracles import port_lib
from parity_cases import rel_err, yarn_cloth_scene

G, P = ptb.load(), port_lib()
for n, seg, radius in ((120, 300, 0.003), (80, 150, 0.0045), (40, 80, 0.009), (20, 40, 0.02)):
    for nb in (1, 5):
        a, b = yarn_cloth_scene(P, 512, 512, 1, n, seg, radius), yarn_cloth_scene(G, 512, 512, 1, n, seg, radius)
        a.nb_bounces = b.nb_bounces = nb
        a.commit(); b.commit()
        oa, ob = a.primary_ids(), b.primary_ids()
        ia, ib = a.render_image_nopreviz().copy(), b.render_image_nopreviz().copy()
        bad = rel_err(ia, ib) > 1e-3
        print(f"[cloth] 2x{n} yarns x {seg} segments, radius {radius * 40:.3f} at 50, depth {nb}: ids agree {((oa[0] == ob[0]) & (oa[1] == ob[1])).mean():.6f}  "
              f"pixels off by > 1e-3: {bad.mean():.5f}  mean {ia.mean():.2f} / {ib.mean():.2f} ({ib.mean() / ia.mean() - 1:+.2e})  "
              f"rays closest {a.stats['rays_closest']} / {b.stats['rays_closest']}  shadow {a.stats['rays_shadow']} / {b.stats['rays_shadow']}  gpu ms {b.stats['ms_device']:.2f}", flush=True)
        a.close(); b.close()
