"""torchrun --nproc-per-node N scripts/multi_gpu_check.py : the tile-sharded NCCL render equals the single-GPU render."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch, torch.distributed as dist
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, multi
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = ptb.load()
rt = scenes.config_C2(lib, 700, 500, 8, nv=120, env=(512, 256), device=local).commit()
img, stats = multi.render_sharded(rt, rank, world, dev)
tot = torch.tensor([stats["samples"]], dtype=torch.int64, device=dev); dist.all_reduce(tot)
if rank == 0:
    cnt = rt.sample_count.copy()
    ref = rt.render_image_nopreviz().copy()
    ok = bool(np.allclose(img, ref, rtol=2e-5, atol=1e-3) and np.allclose(cnt, rt.sample_count, rtol=2e-5))
    print(json.dumps({"world": world, "samples_all_ranks": int(tot[0]), "expected": 700 * 500 * 8, "max_abs_diff": float(np.abs(img - ref).max()),
                      "mean": float(ref.mean()), "sharded_equals_single": ok}), flush=True)
    assert ok and int(tot[0]) == 700 * 500 * 8
dist.barrier(); dist.destroy_process_group()
