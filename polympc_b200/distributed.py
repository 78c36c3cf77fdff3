"""Multi-GPU plumbing: one process per GPU, instances sharded in contiguous blocks, NO collective on the solve path.

SURVEY.md §8(e): the instances of a batch are fully independent (nothing couples them anywhere in SQPBase::solve), so rank r
owns rows [lo, hi) of the batch and its slice of every state array.  The only collectives are (i) one broadcast of the
shared Chebyshev tables at init (checked against the locally computed tables) and (ii) optional gathers of results /
timings.  torch.distributed is plumbing only (NCCL on GPUs, gloo in the CPU test-suite).
"""
from __future__ import annotations

import numpy as np

from . import workloads as W


def _dist():
    import torch.distributed as dist
    return dist


def world():
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def broadcast_tables(api, P: int, device="cpu"):
    """Rank 0 broadcasts the Chebyshev nodes / differentiation matrix / quadrature weights; every rank checks that its own
    (deterministically computed) tables are bit-identical.  Returns the tables."""
    import torch
    nodes, D, w = api.cheb_tables(P)
    flat = torch.tensor(np.concatenate([nodes, D.ravel(), w]), dtype=torch.float64, device=device)
    rank, n = world()
    if n > 1:
        ref = flat.clone()
        _dist().broadcast(ref, src=0)
        if not torch.equal(ref, flat):
            raise RuntimeError(f"rank {rank}: Chebyshev tables differ from rank 0")
    return nodes, D, w


def solve_sharded(api, w: W.Workload, device: int = 0):
    """Solve this rank's shard of workload `w`; returns (lo, hi, solver) with the solver holding the shard's results."""
    rank, n = world()
    lo, hi = W.shard_bounds(w.batch, n, rank)
    if hi <= lo:
        return lo, hi, None
    s = api.sqp(w.name, hi - lo, device)
    W.configure(s, w, lo, hi)
    s.solve()
    return lo, hi, s


def gather_rows(local: np.ndarray, lo: int, hi: int, total: int, device="cpu"):
    """all-gather per-instance rows (any trailing shape) into the full batch order; every rank gets the full array"""
    import torch
    rank, n = world()
    if n == 1:
        return local
    dist = _dist()
    per = -(-total // n)
    pad = np.zeros((per,) + local.shape[1:], dtype=local.dtype)
    pad[:hi - lo] = local
    t = torch.from_numpy(pad).to(device)
    out = [torch.empty_like(t) for _ in range(n)]
    dist.all_gather(out, t)
    full = np.concatenate([o.cpu().numpy() for o in out], axis=0)[:total]
    return full


def max_over_ranks(v: float, device="cpu") -> float:
    import torch
    rank, n = world()
    if n == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=device)
    _dist().all_reduce(t, op=_dist().ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(v: float, device="cpu") -> float:
    import torch
    rank, n = world()
    if n == 1:
        return v
    t = torch.tensor([v], dtype=torch.float64, device=device)
    _dist().all_reduce(t, op=_dist().ReduceOp.SUM)
    return float(t.item())
