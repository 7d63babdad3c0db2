#!/bin/bash
# visit m5: k_shade<MERL> register variants with the out-of-line double fallback; triangle-phase threshold
set -x
mkdir -p gpurun_out
for mb in 5 6 8; do
  PTB_SHADE_MINB_MERL=$mb timeout 300 python scripts/gpu_ab2.py merl_minb$mb C4:256 2>&1 | tee -a gpurun_out/m5_ab.log
done
timeout 600 python - <<'P' 2>&1 | tee gpurun_out/m5_tri_pct.log
import sys; sys.path.insert(0, '.')
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, _abi
G = ptb.load()
for wl, spp in (("C2", 128), ("C3", 64), ("C4", 128)):
    rt = scenes.CONFIGS[wl](G); rt.nrays = spp; rt.commit()
    for pct in (25, 0, 25, 35, 50, 65):
        rt.set_option(_abi.OPT_TRI_MIN_PCT, pct)
        best = min((rt.render_image_nopreviz(want_image=False), rt.stats["ms_device"])[1] for _ in range(3))
        print(f"{wl} spp={spp} tri_min_pct={pct}: {best:8.2f} ms", flush=True)
    rt.close()
P
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/m5_pytest.log
