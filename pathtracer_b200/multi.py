"""Tile-sharded render over the GPUs of one box: one process per GPU, one NCCL gather at the end.

The frame is cut into tiles; rank r renders the tiles with `tile_id % world == r` (tile ids count row by row, each row rotated
against the one above so that no rank gets a vertical stripe, ptb_scene.h `shard_tile_shift`) into its own
full-frame float4 accumulator (only its tiles plus a ceil(2*sigma) apron are non-zero), packs those
tiles+aprons densely, and rank 0 gathers the packs over NCCL/NVLink, adds them into its accumulator
and resolves.  There is no other exchange: scene and BVH are replicated.

torch is plumbing here (device buffers + torch.distributed); all arithmetic is in the C-ABI library.
The same function runs on CPU tensors with the `gloo` backend against tests/devsim, which is how the
host-side logic is covered without GPUs.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _abi


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def render_sharded(rt, rank, world, device, tile_size=0, want_image=True, group=None):
    """Render `rt`'s frame cooperatively.  Returns (imagedouble, stats) on rank 0 and (None, stats) elsewhere.
    `rt` must be committed on this rank's device."""
    L, ctx = rt.lib, rt._ctx
    W, H = rt.W, rt.H
    rgbw = torch.zeros(H * W * 4, dtype=torch.float32, device=device)
    stats = rt.render_accum(rgbw.data_ptr(), rank, world, tile_size)
    if world == 1:
        return rt.resolve(rgbw.data_ptr(), want_image), stats
    sizes = []
    for r in range(world):
        n = C.c_int64(0)
        p = rt.params(r, world, tile_size)
        L.check(L.shard_pack_size(C.byref(p), r, C.byref(n)), ctx)
        sizes.append(n.value)
    nmax = max(max(sizes), 4)
    packed = torch.zeros(nmax, dtype=torch.float32, device=device)
    p = rt.params(rank, world, tile_size)
    if sizes[rank] and rank != 0:
        L.check(L.shard_pack(ctx, C.byref(p), rank, _ptr(rgbw), _ptr(packed)), ctx)
    if rank == 0:
        bufs = [torch.empty(nmax, dtype=torch.float32, device=device) for _ in range(world)]
        dist.gather(packed, gather_list=bufs, dst=0, group=group)
        for r in range(1, world):
            if sizes[r]:
                pr = rt.params(r, world, tile_size)
                L.check(L.shard_unpack_add(ctx, C.byref(pr), r, _ptr(bufs[r]), _ptr(rgbw)), ctx)
        return rt.resolve(rgbw.data_ptr(), want_image), stats
    dist.gather(packed, gather_list=None, dst=0, group=group)
    return None, stats
