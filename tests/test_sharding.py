"""Host-side multi-rank logic on CPU: world_size-2 gloo (the N>1 path of bench.py / register_sharded)."""
import os
import socket

import numpy as np
import pytest


def _fake_register(srcs, dsts):
    """Deterministic stand-in for the GPU call: the 'transform' encodes the inputs."""
    out = []
    for s, d in zip(srcs, dsts):
        T = np.eye(4)
        T[0, 3] = float(s.sum())
        T[1, 3] = float(d.sum())
        out.append(dict(T=T, fitness=float(len(s)) / 7.0, converged=True, valid=len(d) % 2 == 0, iterations=3,
                        n_linearize=4, n_error=4, lm_failed=False))
    return out


def _make(n):
    rng = np.random.default_rng(9)
    srcs = [rng.normal(size=(10 + i, 3)).astype(np.float32) for i in range(n)]
    dsts = [rng.normal(size=(20 + 2 * i, 3)).astype(np.float32) for i in range(n)]
    return srcs, dsts


def test_shard_pairs_partitions_exactly():
    from b200reg.sharding import shard_pairs
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 4, 8):
            got = sorted(i for r in range(world) for i in shard_pairs(n, world, r))
            assert got == list(range(n))
            cap = (n + world - 1) // world
            assert all(len(shard_pairs(n, world, r)) <= cap for r in range(world))
    costs = [1000 * (1 + (i * 7919) % 13) for i in range(100)]
    for world in (2, 4, 8):
        bins = [shard_pairs(100, world, r, costs) for r in range(world)]
        assert sorted(i for b in bins for i in b) == list(range(100))
        loads = [sum(costs[i] for i in b) for b in bins]
        assert max(loads) - min(loads) <= max(costs)  # LPT keeps the spread within one item
    assert shard_pairs(512, 8, 3) == list(range(192, 256))  # config 4: static 64 per GPU


def test_single_rank_roundtrip():
    from b200reg.sharding import register_sharded
    srcs, dsts = _make(5)
    res = register_sharded(_fake_register, srcs, dsts)
    ref = _fake_register(srcs, dsts)
    for a, b in zip(res, ref):
        assert np.array_equal(a["T"], b["T"]) and a["fitness"] == b["fitness"] and a["valid"] == b["valid"]


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200reg.sharding import register_sharded
    srcs, dsts = _make(n)
    costs = [len(s) + len(d) for s, d in zip(srcs, dsts)]
    res_static = register_sharded(_fake_register, srcs, dsts, dist=dist)
    res_lpt = register_sharded(_fake_register, srcs, dsts, dist=dist, costs=costs)
    q.put((rank, [r["T"].tobytes() + np.float64(r["fitness"]).tobytes() for r in res_static],
           [r["T"].tobytes() for r in res_lpt]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo_gather_is_identical_to_single_rank():
    import torch.multiprocessing as mp
    import sys
    REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = os.pathsep.join([os.path.join(REPO, "fast-lio-sam-qn_b200"), REPO, os.environ.get("PYTHONPATH", "")])
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n = 7  # odd: the last rank gets a padded slot
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=100) for _ in range(2)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    from b200reg.sharding import register_sharded
    srcs, dsts = _make(n)
    single = register_sharded(_fake_register, srcs, dsts)
    want = [r["T"].tobytes() + np.float64(r["fitness"]).tobytes() for r in single]
    for rank, stat, lpt in got:
        assert stat == want, "gathered bytes must not depend on the world size"
        assert lpt == [r["T"].tobytes() for r in single]
