"""Marker-sharded multi-GPU execution: one process per GPU (torchrun), torch.distributed for the
plumbing.

  - markers (columns of M = rows of Mt) are split into contiguous, 128-aligned shards;
  - every rank decodes and contracts its own slice: partial C_g = M_g M_g^T (int32, exact);
  - the ONE exchange on the path: all-reduce(sum, int32) of the n x n partial (NCCL over NVLink);
  - scan outputs stay sharded; the per-step outlier pick is an all-gather of (max tsq, index)
    pairs combined with the reference's tie rule (first index of the maximum, R/find_qtl.R:76-80).

The shard arithmetic and the argmax combine are plain host logic and are tested on CPU with the
gloo backend (tests/test_dist_cpu.py); the kernels are libeaglegpu's.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def shard_range(L: int, world: int, rank: int, align: int = 128):
    """Contiguous marker range [c0, c1) of `rank`: shards are multiples of `align` markers (the K tile
    of the int8 contraction) except the last one, and cover [0, L) exactly."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    blocks = (L + align - 1) // align
    base, extra = divmod(blocks, world)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    return min(b0 * align, L), min(b1 * align, L)


def combine_argmax(best: torch.Tensor, idx: torch.Tensor):
    """best[r], idx[r]: per-rank maximum of tsq (NaN already ignored) and its GLOBAL marker index
    (idx < 0: the shard had no finite/inf value).  Returns (max, lowest index attaining it)."""
    best = best.reshape(-1).to(torch.float64)
    idx = idx.reshape(-1).to(torch.int64)
    win_v, win_i = float("nan"), -1
    for v, i in zip(best.tolist(), idx.tolist()):
        if i < 0 or math.isnan(v):
            continue
        if win_i < 0 or v > win_v or (v == win_v and i < win_i):
            win_v, win_i = v, i
    return win_v, win_i


def allreduce_partial_mmt(C32: torch.Tensor, group=None):
    """The path's single collective: exact int32 sum of the per-shard n x n partial products."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(C32, op=dist.ReduceOp.SUM, group=group)
    return C32


def global_argmax(best: torch.Tensor, idx_local: torch.Tensor, marker_offset: int, group=None):
    """all-gather of 16 bytes per rank, then the reference's first-maximum rule."""
    gidx = torch.where(idx_local >= 0, idx_local + marker_offset, idx_local)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        w = dist.get_world_size(group)
        mine = torch.cat([best.reshape(1).view(torch.int64), gidx.reshape(1).to(torch.int64)])  # (tsq bits, index)
        allr = torch.empty(2 * w, dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(allr, mine, group=group)                                   # ONE collective, one D2H
        allr = allr.cpu().view(w, 2)
        return combine_argmax(allr[:, 0].contiguous().view(torch.float64), allr[:, 1].contiguous())
    return combine_argmax(best.cpu(), gidx.cpu())


def gather_sharded(vec: torch.Tensor, L: int, world: int, group=None):
    """Concatenate per-rank marker vectors (a or vara) into the full length-L vector on every rank
    (the R API returns full vectors)."""
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return vec
    sizes = [shard_range(L, world, r)[1] - shard_range(L, world, r)[0] for r in range(world)]
    m = max(sizes)
    pad = torch.zeros(m, dtype=vec.dtype, device=vec.device)
    pad[: vec.numel()] = vec
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)])


class Shard:
    """This rank's marker shard of a data set of L markers, and the three exchanges a sharded forward search needs
    (am.AM_resident): the int32 all-reduce of the partial M.Mt, the sharded first-maximum pick, and the broadcast of a
    picked marker's genotype column by the rank that owns it."""

    def __init__(self, L: int, world: int, rank: int, group=None):
        self.L, self.world, self.rank, self.group = L, world, rank, group
        self.ranges = [shard_range(L, world, r) for r in range(world)]
        self.c0, self.c1 = self.ranges[rank]

    def allreduce_mmt(self, C32):
        return allreduce_partial_mmt(C32, self.group)

    def global_argmax(self, best, idx_local):
        return global_argmax(best, idx_local, self.c0, self.group)

    def owner(self, g: int) -> int:
        for r, (a, b) in enumerate(self.ranges):
            if a <= g < b:
                return r
        raise ValueError(f"marker {g} outside [0, {self.L})")

    def fetch_col(self, extract_local, n: int, g: int, device):
        """extract_local(j) -> int32[n] device tensor of LOCAL marker j; every rank receives marker g's column."""
        r = self.owner(g)
        col = extract_local(g - self.c0) if r == self.rank else torch.empty(n, dtype=torch.int32, device=device)
        if self.world > 1:
            dist.broadcast(col, src=r, group=self.group)
        return col
