"""N>1 path on CPU: two processes, `gloo` backend, the sharded render + tile gather of pathtracer_b200/multi.py,
with tests/devsim standing in for the CUDA library (same ABI, host memory instead of device memory)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port_no, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracles import devsim_lib
    from pathtracer_b200 import multi, scenes
    lib = devsim_lib()
    rt = scenes.config_C2(lib, 150, 70, 3, nv=16, env=(64, 32)).commit()
    img, stats = multi.render_sharded(rt, rank, world, torch.device("cpu"), tile_size=32)
    samples = torch.tensor([stats["samples"]], dtype=torch.int64)
    dist.all_reduce(samples)
    if rank == 0:
        np.savez(out_path, img=img, cnt=rt.sample_count, samples=samples.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_render_equals_single_process(tmp_path, world, devsim):
    from pathtracer_b200 import scenes
    out = str(tmp_path / "out.npz")
    port_no = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port_no, out), nprocs=world, join=True)
    got = np.load(out)
    rt = scenes.config_C2(devsim, 150, 70, 3, nv=16, env=(64, 32)).commit()
    ref = rt.render_image_nopreviz()
    assert got["samples"][0] == 150 * 70 * 3
    assert np.allclose(got["img"], ref, rtol=2e-5, atol=1e-3) and np.allclose(got["cnt"], rt.sample_count, rtol=2e-5)
