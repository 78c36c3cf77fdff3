"""N > 1 path on CPU: world_size-2 gloo run of the sharding host logic (polympc_b200/distributed.py).  The solver behind the
C ABI is the oracle here (no GPU in this suite); the same code drives libpolympc_b200 under NCCL in bench.py."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT
from polympc_b200 import workloads as W


def test_shard_bounds_cover_the_batch():
    for batch in (1, 7, 8, 8192, 1000):
        for n in (1, 2, 3, 4, 8):
            cuts = [W.shard_bounds(batch, n, r) for r in range(n)]
            assert cuts[0][0] == 0 and cuts[-1][1] == batch
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0]
            assert all(hi - lo <= -(-batch // n) for lo, hi in cuts)


def _worker(rank, world_size, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    import torch.distributed as dist
    from oracle import pyoracle
    from polympc_b200 import distributed as D
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    api = pyoracle.load()
    D.broadcast_tables(api, 6)
    w = W.mobile_robot(9, seed=5, sqp_max_iter=6, ls_max_iter=10)     # 9 instances over 2 ranks: shards of 5 and 4
    lo, hi, s = D.solve_sharded(api, w)
    x = D.gather_rows(s.primal(), lo, hi, w.batch)
    it = D.gather_rows(s.info()["iter"].astype(np.int64), lo, hi, w.batch)
    total = D.sum_over_ranks(float(s.info()["iter"].sum()))
    tmax = D.max_over_ranks(float(rank + 1))
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), x=x, it=it, total=total, tmax=tmax)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_solve_matches_single_process(tmp_path, orc):
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "sharded.npz")
    w = W.mobile_robot(9, seed=5, sqp_max_iter=6, ls_max_iter=10)
    s = orc.sqp(w.name, 9)
    W.configure(s, w)
    s.solve()
    assert np.array_equal(got["x"], s.primal())                      # sharding does not change any instance's result
    assert np.array_equal(got["it"], s.info()["iter"])
    assert float(got["total"]) == float(s.info()["iter"].sum())
    assert float(got["tmax"]) == 2.0
