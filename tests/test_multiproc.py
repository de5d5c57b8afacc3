"""CPU tier: the N > 1 sharding path on gloo, world_size 2 (no GPU needed: the per-shard compute is
the oracle here, the point is the partitioning + final gather logic bench.py / the engine rely on)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from plonky2_ecdsa_b200 import sharding, workload as W
    from tests import util as U
    x, y = W.mulmod_workload(n, 0, seed=11)          # every rank derives the same global batch from the seed

    def compute(lo, hi):
        return U.expect_mul(x[:, lo:hi], y[:, lo:hi], 0)
    full, err_all, (lo, hi) = sharding.run_sharded(n, compute, dist=dist, gather_results=True)
    local, err2, _ = sharding.run_sharded(n, compute, dist=dist, gather_results=False)
    # the bench's strong-scaling gather: column shards of a uint32 matrix into rank 0
    g = sharding.RootGather(n, 51, torch.int64, torch.device("cpu"), dist, dst=0)
    g(torch.from_numpy(local.view(np.int64).copy()))
    g(torch.from_numpy(local.view(np.int64).copy()))          # buffers are reusable
    rooted = g.assemble().numpy().view(np.uint64) if rank == 0 else None
    q.put((rank, lo, hi, full, err_all, local, err2, rooted))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 37])
def test_world_size_2_sharding_and_gather(n):
    from plonky2_ecdsa_b200 import sharding, workload as W
    from tests import util as U
    assert sharding.shard_range(37, 0, 2) == (0, 18) and sharding.shard_range(37, 1, 2) == (18, 37)
    assert sharding.shard_range(5, 3, 8) == (1, 2) and sharding.shard_range(0, 0, 2) == (0, 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x, y = W.mulmod_workload(n, 0, seed=11)
    eo, ee = U.expect_mul(x, y, 0)
    cover = []
    for rank, lo, hi, full, err_all, local, err2, rooted in res:
        assert np.array_equal(full, eo) and np.array_equal(err_all, ee) and np.array_equal(err2, ee)
        assert (rooted is None) == (rank != 0)
        if rooted is not None:
            assert np.array_equal(rooted, eo)
        assert np.array_equal(local, eo[:, lo:hi])
        cover.append((lo, hi))
    assert sorted(cover) == [(0, n // 2), (n // 2, n)]
