"""First GPU contact: parity of the CUDA path vs oracle/_ref on small scenes + raw throughput of C2/C3."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, _abi
from oracles import ref_lib

G = ptb.load()
R = ref_lib()
print("version", G.version().decode(), "| ref:", R.version().decode() if R else None, flush=True)

def parity(name, mk):
    rg = mk(G).commit()
    og, tg, dg = rg.primary_ids()
    t = time.time(); b = rg.render_image_nopreviz().copy(); tgpu = time.time() - t
    line = {"scene": name, "gpu_mean": float(b.mean()), "gpu_s": tgpu, "gpu_stats": rg.stats}
    if R is not None:
        rr = mk(R).commit()
        oa, ta, da = rr.primary_ids()
        a = rr.render_image_nopreviz().copy()
        rel = np.abs(a - b) / np.maximum(np.abs(a), 1e-3 * a.mean())
        line.update(ref_mean=float(a.mean()), id_agree=float(np.mean((oa == og) & (ta == tg))),
                    frac_rel_gt_1e3=float(np.mean(rel.max(-1) > 1e-3)), frac_rel_gt_1e2=float(np.mean(rel.max(-1) > 1e-2)),
                    ref_rays=[rr.stats["rays_closest"], rr.stats["rays_shadow"]])
    print(json.dumps(line), flush=True)
    rg.close()

parity("C1s", lambda L: scenes.config_C1(L, 128, 128, 4))
parity("C2s", lambda L: scenes.config_C2(L, 128, 128, 2, nv=40, env=(256, 128)))
parity("C3s", lambda L: scenes.config_C3(L, 128, 128, 2, nv=40, tex=128))
parity("C4s", lambda L: scenes.config_C4(L, 128, 128, 2, nv=40))
parity("C5s", lambda L: scenes.config_C5(L, 160, 96, 1, nv=20))

def perf(name, mk, reps=2, count=False):
    t = time.time(); rt = mk(G); tgen = time.time() - t
    t = time.time(); rt.commit(); tcommit = time.time() - t
    info = rt.scene_info()
    if count:
        rt.set_option(_abi.OPT_COUNT_TRAVERSAL, 1)
    for r in range(reps):
        rt.render_image_nopreviz(want_image=False)
        s = rt.stats
        rays = s["rays_closest"] + s["rays_shadow"]
        print(json.dumps({"scene": name, "rep": r, "count": count, "gen_s": tgen, "commit_s": tcommit, "bvh_ms": info["ms_bvh_build"], "nodes": info["n_bvh_nodes"],
                          "depth": info["bvh_depth"], "ms_device": s["ms_device"], "ms_wall": s["ms_wall"], "Msamples_s": s["samples"] / s["ms_device"] / 1e3,
                          "Mrays_s": rays / s["ms_device"] / 1e3, "rays_per_sample": rays / s["samples"], "launches": s["kernel_launches"],
                          "nodes_per_ray": s["node_visits"] / rays, "tris_per_ray": s["tri_tests"] / rays, "mean": float(rt.imagedouble.mean())}), flush=True)
    rt.close()

perf("C1", lambda L: scenes.config_C1(L))
perf("C2", lambda L: scenes.config_C2(L, spp=64))
perf("C2", lambda L: scenes.config_C2(L, spp=16), reps=1, count=True)
perf("C3", lambda L: scenes.config_C3(L, spp=32))
perf("C3", lambda L: scenes.config_C3(L, spp=8), reps=1, count=True)
perf("C4", lambda L: scenes.config_C4(L, spp=64))
