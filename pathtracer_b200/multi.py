"""Tile-sharded render over the GPUs of one box: one process per GPU, one NCCL gather at the end.

The frame is cut into tiles; rank r renders the tiles with `tile_id % world == r` (tile ids count row by row, each row rotated
against the one above so that no rank gets a vertical stripe, ptb_scene.h `shard_tile_shift`) into its own
full-frame float4 accumulator (only its tiles plus a ceil(2*sigma) apron are non-zero), packs those
tiles+aprons densely, and rank 0 gathers the packs over NCCL/NVLink, adds them into its accumulator
and resolves.  There is no other exchange: scene and BVH are replicated.

Two routes:
  * `init_comm(rt, rank, world)` + `rt.render_image_nopreviz()`: the product route.  The gather runs INSIDE the library
    (ptb_render_sharded: pack -> ncclSend/ncclRecv -> unpack-add -> resolve on the context's own stream); torch.distributed only
    carries the 128-byte NCCL id from rank 0 to the other ranks.
  * `render_sharded(...)`: the same steps driven from the host through ptb_render_accum / ptb_shard_* and a
    torch.distributed gather.  It runs on CPU tensors with the `gloo` backend against tests/devsim, which is how the tile
    ownership / pack / unpack logic is covered without GPUs.  Stream contract (include/ptb200.h): the library works on its own
    non-blocking stream, so device buffers produced by torch are synchronised before they are handed over.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _abi


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def init_comm(rt, rank, world, group=None):
    """Give `rt` (any state; survives commit) its place in the NCCL communicator of the tile-sharded render."""
    if world == 1:
        return rt.comm_init(1, 0, None)
    ids = [rt.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0, group=group)
    return rt.comm_init(world, rank, ids[0])


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


def render_sharded(rt, rank, world, device, tile_size=0, want_image=True, group=None):
    """Render `rt`'s frame cooperatively.  Returns (imagedouble, stats) on rank 0 and (None, stats) elsewhere.
    `rt` must be committed on this rank's device."""
    L, ctx = rt.lib, rt._ctx
    W, H = rt.W, rt.H
    rgbw = torch.zeros(H * W * 4, dtype=torch.float32, device=device)
    _sync(device)                # the fill ran on torch's stream, the library accumulates on its own
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
    _sync(device)
    p = rt.params(rank, world, tile_size)
    if sizes[rank] and rank != 0:
        L.check(L.shard_pack(ctx, C.byref(p), rank, _ptr(rgbw), _ptr(packed)), ctx)
    if rank == 0:
        bufs = [torch.empty(nmax, dtype=torch.float32, device=device) for _ in range(world)]
        dist.gather(packed, gather_list=bufs, dst=0, group=group)
        _sync(device)            # a blocking NCCL call only orders torch's stream: wait until the receives have landed
        for r in range(1, world):
            if sizes[r]:
                pr = rt.params(r, world, tile_size)
                L.check(L.shard_unpack_add(ctx, C.byref(pr), r, _ptr(bufs[r]), _ptr(rgbw)), ctx)
        return rt.resolve(rgbw.data_ptr(), want_image), stats
    dist.gather(packed, gather_list=None, dst=0, group=group)
    return None, stats
