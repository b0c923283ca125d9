"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on the box, gloo in
the CPU tests).  GFPush sources and aggregation batches are independent, so the data path has NO
collective: each rank pushes its contiguous shard of sources on a replicated CSR; the only exchange
is an all-gather of the finished [S_g, K] (col, val) rows when every rank needs the whole Pi."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of `total` items for `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_gather_rows(shards: Sequence[torch.Tensor], total_rows: int, group=None) -> List[torch.Tensor]:
    """All-gather row-sharded tensors ([S_g, ...] each, shards laid out by shard_range) into full
    [total_rows, ...] tensors on every rank.  Ragged shards are padded to the largest one."""
    rank, ws = world()
    if ws == 1:
        return list(shards)
    out = []
    sizes = [shard_range(total_rows, r, ws) for r in range(ws)]
    max_rows = max(hi - lo for lo, hi in sizes)
    for t in shards:
        pad = torch.zeros((max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        buf = torch.empty((ws * max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, pad, group=group)
        buf = buf.reshape((ws, max_rows) + tuple(t.shape[1:]))
        out.append(torch.cat([buf[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], 0))
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over ranks (timing: a multi-GPU step takes as long as its slowest rank)."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    rank, ws = world()
    if ws == 1:
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gfpush_sharded(graph, node_idx, coef, rmax, K, gather: bool = True, check: bool = True):
    """GFPush of `node_idx` (identical on every rank) with sources sharded by rank.  Returns the
    device tensors (col int32, val64, val32) for all sources when gather=True (all-gather over
    NCCL/NVLink), else for this rank's shard only, plus the shard bounds."""
    rank, ws = world()
    dev = torch.device("cuda", graph.device)
    nid = torch.as_tensor(node_idx).to(device=dev, dtype=torch.int32)
    lo, hi = shard_range(nid.numel(), rank, ws)
    # check: wait for the shard's push and raise on a device-side refusal before the rows are gathered
    _row, col, val, val32 = graph.gfpush_device(nid[lo:hi].contiguous(), coef, rmax, K, want_fp32=True, check=check)
    if gather and ws > 1:
        col, val, val32 = all_gather_rows([col, val, val32], nid.numel())
    return col, val, val32, (lo, hi)


def allreduce_sparse_rows(rows: torch.Tensor, vals: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sum row-sparse gradients over the ranks (MAG backward): every rank contributes (rows int64 [R_g] distinct,
    vals [R_g, H]); every rank receives the coalesced union (rows ascending, values summed).  The exchange is one all-gather
    of the padded (row, value) lists -- a batch touches a few thousand of the table's 2.78 M rows, so this moves kilobytes
    where the dense all-reduce of /root/reference/model_mag.py:27's gradient would move 713 MB."""
    rank, ws = world()
    if ws == 1:
        return rows, vals
    dev = vals.device
    n = torch.tensor([rows.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    H = int(vals.shape[1])
    prow = torch.full((m,), -1, dtype=torch.int64, device=dev)
    prow[: rows.numel()] = rows
    pval = torch.zeros((m, H), dtype=vals.dtype, device=dev)
    pval[: rows.numel()] = vals
    grow = torch.empty((ws * m,), dtype=torch.int64, device=dev)
    gval = torch.empty((ws * m, H), dtype=vals.dtype, device=dev)
    dist.all_gather_into_tensor(grow, prow, group=group)
    dist.all_gather_into_tensor(gval, pval, group=group)
    keep = grow >= 0
    grow, gval = grow[keep], gval[keep]
    urows, inv = torch.unique(grow, return_inverse=True)
    out = torch.zeros((urows.numel(), H), dtype=vals.dtype, device=dev)
    out.index_add_(0, inv, gval)     # rank order is fixed, so every rank sums in the same order
    return urows, out
