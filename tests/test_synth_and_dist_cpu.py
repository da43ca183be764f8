"""CPU tests of the host logic around the kernels: synthetic generator, marker sharding, the
argmax tie rule, and the N>1 path over the gloo backend with world_size 2."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eagleeverything_b200 import dist as egd
from eagleeverything_b200 import synth


def test_synth_is_deterministic_and_shard_consistent():
    G = synth.genotypes(64, 700)
    assert G.dtype == np.uint8 and set(np.unique(G)) <= {0, 1, 2}
    assert np.array_equal(G, synth.genotypes(64, 700))
    assert np.array_equal(G[:, 300:555], synth.genotypes(64, 255, col_offset=300, n_total=64))
    assert np.array_equal(G[10:20], synth.genotypes(10, 700, n_total=64, row_offset=10))
    img = synth.ascii_image(G)
    assert img.shape == (64, 701) and (img[:, -1] == 10).all() and img[:, :-1].min() >= 48


@pytest.mark.parametrize("L,world", [(1000000, 8), (4998, 2), (127, 4), (128, 4), (129, 2), (1, 8), (600000, 3)])
def test_shard_range_covers_exactly(L, world):
    prev = 0
    for r in range(world):
        c0, c1 = egd.shard_range(L, world, r)
        assert c0 == prev and c0 <= c1 <= L
        if c1 < L:
            assert c1 % 128 == 0
        prev = c1
    assert prev == L
    sizes = [egd.shard_range(L, world, r)[1] - egd.shard_range(L, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 256 or L < 128 * world


def test_combine_argmax_tie_rule():
    # ties -> lowest global index, regardless of which rank holds it (R/find_qtl.R:76-80)
    v, i = egd.combine_argmax(torch.tensor([5.0, 7.0, 7.0]), torch.tensor([10, 900, 400]))
    assert (v, i) == (7.0, 400)
    v, i = egd.combine_argmax(torch.tensor([float("nan"), 3.0]), torch.tensor([-1, 77]))
    assert (v, i) == (3.0, 77)
    v, i = egd.combine_argmax(torch.tensor([float("inf"), 3.0]), torch.tensor([5, 2]))
    assert v == float("inf") and i == 5
    v, i = egd.combine_argmax(torch.tensor([float("nan")]), torch.tensor([-1]))
    assert i == -1 and np.isnan(v)


def _worker(rank, world, port, n, L, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        G = synth.genotypes(n, L).astype(np.int32) - 1
        c0, c1 = egd.shard_range(L, world, rank)
        part = torch.from_numpy((G[:, c0:c1] @ G[:, c0:c1].T).astype(np.int32))  # stands in for the SYRK kernel
        full = egd.allreduce_partial_mmt(part.clone())
        ok_mmt = bool(np.array_equal(full.numpy(), G @ G.T))
        # sharded scores with a tie across ranks: marker 3 (rank 0) and marker L-2 (rank 1) identical
        rng = np.random.default_rng(0)
        tsq = rng.random(L)
        tsq[3] = tsq[L - 2] = 2.0
        tsq[5] = np.nan
        loc = tsq[c0:c1]
        k = int(np.nanargmax(loc))
        best, idx = egd.global_argmax(torch.tensor([loc[k]]), torch.tensor([k]), c0)
        vec = egd.gather_sharded(torch.from_numpy(loc.copy()), L, world)
        ok_gather = bool(np.array_equal(np.nan_to_num(vec.numpy(), nan=-1), np.nan_to_num(tsq, nan=-1)))
        # dist.Shard: what the sharded forward search uses (owner of a marker, its column broadcast, the sharded pick)
        sh = egd.Shard(L, world, rank)
        cols = {}
        first_end = egd.shard_range(L, world, 0)[1]
        for gidx in (0, first_end - 1, first_end, L - 1, 128):          # the same markers on every rank (collective)
            col = sh.fetch_col(lambda j: torch.from_numpy(G[:, c0 + j].astype(np.int32)), n, gidx, "cpu")
            cols[gidx] = bool(np.array_equal(col.numpy(), G[:, gidx].astype(np.int32)))
        owners = [sh.owner(0), sh.owner(L - 1)]
        b2, i2 = sh.global_argmax(torch.tensor([loc[k]]), torch.tensor([k]))
        ret[rank] = (ok_mmt, best, idx, ok_gather, all(cols.values()), owners, b2, i2)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    world, n, L = 2, 37, 1000
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, L, ret), nprocs=world, join=True)
    for r in range(world):
        ok_mmt, best, idx, ok_gather, ok_cols, owners, b2, i2 = ret[r]
        assert ok_mmt and ok_gather and ok_cols
        assert best == 2.0 and idx == 3 and b2 == 2.0 and i2 == 3
        assert owners == [0, world - 1]
