"""Multi-GPU plumbing for the sharded LBG train: one process per GPU, block rows
partitioned across ranks, one integer all-reduce of K*(dim+2) 64-bit words per split level.

The data path itself lives in libqb200; this module only supplies
  * ``shard_rows``      - which block rows (and which image bytes) a rank owns,
  * ``make_allreduce``  - the callback handed to ``qb200_train`` (torch.distributed, NCCL on GPUs,
                          gloo on CPU for the host-logic tests).
Integer sums make the reduced statistics - and therefore every rank's next codebook - identical
bit for bit whatever the rank count or reduction order.
"""
from __future__ import annotations

from typing import Tuple


def shard_rows(w_blocks: int, world: int, rank: int) -> Tuple[int, int]:
    """Block rows [begin, end) owned by ``rank``: contiguous, balanced to within one row."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(w_blocks, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_byte_range(xSize: int, ySize: int, w: int, h: int, row_begin: int, row_end: int) -> Tuple[int, int]:
    """Image bytes [lo, hi) a shard reads (mirrors qb200_set_image_shard): a block row is the
    contiguous range of w*ySize pixels; with ySize % h != 0 the last block of a row spills h-1
    pixels at most into the following bytes (the reference's y-overflow wrap)."""
    img_bytes = xSize * ySize * 3
    h_blocks = (ySize + h - 1) // h
    lo = min(row_begin * w * ySize * 3, img_bytes)
    if row_end <= row_begin:
        return lo, lo
    last_elem = ((w - 1) * ySize + (h - 1)) * 3 + 2
    hi = (row_end - 1) * w * ySize * 3 + (h_blocks - 1) * h * 3 + last_elem + 1
    return lo, max(lo, min(hi, img_bytes))


def join_peer_group(ctx, group=None, max_words: int = 0):
    """One process per GPU: joins this rank's context to the library's own peer-memory all-reduce group
    (qb200_comm_export / qb200_comm_attach; include/qb200.h).  torch.distributed only carries the 64-byte CUDA IPC
    handles once; after this ``ctx.train(...)`` needs no callback - the per-level sum all-reduce runs inside
    libqb200 over NVLink peer memory."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    mine = ctx.comm_export(max_words)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    handles = b"".join(bytes(o.cpu().numpy().tobytes()) for o in out)
    ctx.comm_attach(world, rank, handles)
    dist.barrier(group=group)          # nobody publishes before everybody has mapped everybody
    return world, rank


def make_allreduce(group=None):
    """Returns ``fn(dev_ptr, count, stream) -> int`` summing ``count`` uint64 words in place across
    the ranks of ``group`` with torch.distributed.  The words are viewed as int64 (two's complement
    sum == unsigned sum)."""
    import torch
    import torch.distributed as dist

    def allreduce(dev_ptr: int, count: int, stream: int) -> int:
        t = _wrap_device_int64(dev_ptr, count)
        # enqueue on the library's stream: torch orders the collective after the work already on it and makes
        # the stream wait for the result, which is all qb200_allreduce_fn asks for (no host synchronisation)
        ext = torch.cuda.ExternalStream(stream) if stream else torch.cuda.current_stream()
        with torch.cuda.stream(ext):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return 0

    return allreduce


def _wrap_device_int64(dev_ptr: int, count: int):
    """Zero-copy torch view of ``count`` int64 words at a raw device address."""
    import torch

    class _Holder:
        __cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (dev_ptr, False),
                                    "version": 3, "strides": None}

    return torch.as_tensor(_Holder(), device=torch.device("cuda", torch.cuda.current_device()))


def allreduce_host_int64(arr, group=None):
    """Host-side (gloo) variant used by the CPU tests: sums a numpy int64/uint64 array in place."""
    import numpy as np
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    arr[...] = t.numpy().view(arr.dtype).reshape(arr.shape)
    return arr
