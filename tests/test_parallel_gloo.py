"""World-size-2 gloo test (CPU) of the multi-GPU host logic: round-robin unit sharding + the single
all-reduce of [ELBO, d/d global params] reproduce the unsharded sums."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vbsky_b200.parallel import GLOBAL_KEYS, allreduce_elbo, pack_global, shard_units, unpack_global

T, K, M = 7, 3, 5


def _fake_unit(t, k):
    """Deterministic stand-in for one unit's (ll, global-parameter gradients)."""
    rng = np.random.default_rng(1000 * t + k)
    return rng.normal(), {"origin": rng.normal(), "clock_rate": rng.normal(), "R": rng.normal(size=M),
                          "delta": rng.normal(size=M), "s": rng.normal(size=M), "rho": rng.normal(size=M)}


def _local(units):
    ll = torch.tensor([_fake_unit(t, k)[0] for t, k in units], dtype=torch.float64)
    grads = {}
    for key in GLOBAL_KEYS:
        rows = [np.atleast_1d(_fake_unit(t, k)[1][key]) for t, k in units]
        g = torch.tensor(np.stack(rows), dtype=torch.float64)
        grads[key] = g[:, 0] if key in ("origin", "clock_rate") else g
    return ll, grads


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = shard_units(T, K, rank, world)
    vec = allreduce_elbo(pack_global(*_local(units), K))
    q.put((rank, len(units), vec.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_units_partition():
    for world in (1, 2, 3, 8):
        seen = sorted(u for r in range(world) for u in shard_units(100, 10, r, world))
        assert seen == [(t, k) for t in range(100) for k in range(10)]
        sizes = [len(shard_units(100, 10, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_allreduce_matches_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = pack_global(*_local(shard_units(T, K, 0, 1)), K).numpy()
    assert sum(r[1] for r in res) == T * K
    for _, _, got in res:
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13)
    un = unpack_global(torch.tensor(want), M)
    assert un["R"].shape == (M,) and un["rho"].shape == (M,)
