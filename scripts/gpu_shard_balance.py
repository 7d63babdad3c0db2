"""One GPU renders every shard of an N-way split in turn: per-shard device time (load balance of the tile ownership) and the effect
of the pass pipelines on a shard-sized job.  usage: python scripts/gpu_shard_balance.py [C2] [8]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pathtracer_b200 as ptb
from pathtracer_b200 import scenes, _abi
wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
G = ptb.load()
rt = scenes.CONFIGS[wl](G); rt.commit()
rgbw = torch.zeros(rt.H * rt.W * 4, dtype=torch.float32, device="cuda")
for N in [int(a) for a in sys.argv[2:]] or [8, 4, 2]:
    combos = [(2, int(t)) for t in os.environ["PTB_TILES"].split(",")] if os.environ.get("PTB_TILES") else ((1, 64), (2, 64), (2, 32))
    for pipes, tile in combos:
        rt.set_option(_abi.OPT_PIPES, pipes)
        times = []
        for r in range(N):
            best = 1e30
            for rep in range(3):
                rgbw.zero_(); torch.cuda.synchronize()
                st = rt.render_accum(rgbw.data_ptr(), r, N, tile)
                best = min(best, st["ms_device"])
            times.append(best)
        mean = sum(times) / N
        print(f"{wl} N={N} pipes={pipes} tile={tile}: per-shard ms " + " ".join(f"{t:.2f}" for t in times) + f"  max {max(times):.2f} mean {mean:.2f} max/mean {max(times) / mean:.3f}", flush=True)
