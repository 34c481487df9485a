"""The N>1 host path on CPU: world_size-2 gloo.  Each rank owns a shard of the block rows, computes
its partial integer statistics (here with the oracle standing in for the kernels), all-reduces the
K*(dim+2) words and finalises - every rank must end with the single-process result, bit for bit."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden
from quant_b200.distributed import shard_byte_range, shard_rows


def test_shard_rows_partition():
    for wb in (1, 2, 7, 50, 2048):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(wb, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == wb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(4, 2, 2)


def test_shard_byte_range_covers_exactly_what_blocks_read():
    from quant_b200.compressor import _block_index_map
    for (xs, ys, w, h) in [(8, 6, 2, 2), (9, 7, 2, 3), (101, 67, 3, 2), (16, 16, 4, 4)]:
        idx = _block_index_map(xs, ys, w, h)
        hB = (ys + h - 1) // h
        wB = (xs + w - 1) // w
        for world in (2, 3):
            for r in range(world):
                b, e = shard_rows(wB, world, r)
                lo, hi = shard_byte_range(xs, ys, w, h, b, e)
                used = idx[b * hB:e * hB].reshape(-1)
                used = used[used < xs * ys]
                if used.size:
                    assert lo <= used.min() * 3 and used.max() * 3 + 3 <= hi
                    assert lo == b * w * ys * 3


def _worker(rank, world, port_file, name, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import quant_b200 as qb
    from oracle.pyoracle import PortLib
    from quant_b200.distributed import allreduce_host_int64, shard_rows
    from conftest import load_golden
    dist.init_process_group("gloo", init_method=f"file://{port_file}", rank=rank, world_size=world)
    P = PortLib()
    g = load_golden(name)
    T = P.blocks_lattice(g.rgb, g.xs, g.ys, g.w, g.h, g.cs).astype(np.int64)
    L = T - 128 if g.cs == 1 else T
    hB = (g.ys + g.h - 1) // g.h
    wB = (g.xs + g.w - 1) // g.w
    b, e = shard_rows(wB, world, rank)
    sl = slice(b * hB, e * hB)
    ok = True
    for i in range(g.nbits):
        l = g.level(i)
        K = l["K"]
        a = l["assign"][sl].astype(np.int64)           # what this rank's kernels would produce
        words = np.zeros((K, g.dim + 2), np.int64)
        np.add.at(words[:, 0], a, 1)
        np.add.at(words[:, 1:1 + g.dim], a, L[sl])
        np.add.at(words[:, g.dim + 1], a, (L[sl] ** 2).sum(1))
        allreduce_host_int64(words)
        post, d0, d1 = qb.finalize_level(g.cs, g.N, words[:, 0].astype(np.uint64), words[:, 1:1 + g.dim],
                                         words[:, g.dim + 1].astype(np.uint64), l["cb_pre"])
        ok &= bool(np.max(np.abs(post - l["cb_post"]) / np.maximum(np.abs(l["cb_post"]), 1e-300)) < 1e-14)
        ok &= abs(d1 - l["d1"]) <= 1e-10 * abs(l["d1"]) and abs(d0 - l["d0"]) <= 1e-10 * abs(l["d0"])
        np.save(os.path.join(out_dir, f"post_{rank}_{i}.npy"), post)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(3)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_statistics_allreduce_gloo(port, tmp_path, world):
    import torch.multiprocessing as mp
    name = "odd_101x67_2x2_n6"
    rdv = str(tmp_path / "rdv")
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, rdv, name, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    g = load_golden(name)
    for i in range(g.nbits):
        ref = np.load(tmp_path / f"post_0_{i}.npy")
        for r in range(1, world):
            assert np.array_equal(ref, np.load(tmp_path / f"post_{r}_{i}.npy"))  # ranks agree bit for bit
