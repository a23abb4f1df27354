"""Multi-GPU host logic on CPU: world_size-2 gloo run of the sharding + timing reduction bench.py uses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import FR3


def test_shard_bounds_cover_exactly():
    from rigidbody_rs_b200.shard import shard_bounds
    for total in (0, 1, 7, 1 << 24, 65536, 1000003):
        for world in (1, 2, 3, 4, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rigidbody_rs_b200.shard import max_over_ranks, shard_bounds, sum_over_ranks
    from oracle.rb_oracle import Oracle
    total = 1001
    lo, hi = shard_bounds(total, world, rank)
    # each rank regenerates its own slice of the counter-based sample and runs the (CPU oracle) dynamics on it:
    # the union over ranks must equal the single-process result, with no data exchanged between ranks.
    o = Oracle.from_urdf(FR3)
    m = o.model
    qs = o.fill(0x5EED0001, 0, m.lower, m.upper, lo, hi - lo)
    dqs = o.fill(0x5EED0001, 1, -m.velocity, m.velocity, lo, hi - lo)
    tau = o.rnea_batch(qs, dqs, np.zeros_like(qs))
    t = max_over_ranks(10.0 + rank)
    n = sum_over_ranks(hi - lo)
    q.put((rank, lo, hi, tau, t, n))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding(built):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    from oracle.rb_oracle import Oracle
    o = Oracle.from_urdf(FR3)
    m = o.model
    qa = o.fill(0x5EED0001, 0, m.lower, m.upper, 0, 1001)
    dqa = o.fill(0x5EED0001, 1, -m.velocity, m.velocity, 0, 1001)
    want = o.rnea_batch(qa, dqa, np.zeros_like(qa))
    got = np.concatenate([r[3] for r in res], axis=1)
    np.testing.assert_array_equal(got, want)
    assert [r[4] for r in res] == [11.0, 11.0]          # max over ranks
    assert [r[5] for r in res] == [1001.0, 1001.0]      # units summed over ranks
