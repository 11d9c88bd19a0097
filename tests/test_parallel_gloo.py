"""World-size-2 gloo test (CPU) of the multi-GPU host logic: round-robin unit sharding + the single all-reduce of
[ELBO, d/d global params, d/d per-tree local params] reproduce the unsharded sums -- in particular the local block,
whose K units per tree land on different ranks.  The payload itself is built by a CUDA kernel in the product
(``parallel.reduce_elbo``; checked against ``pack_reference`` below in tests/test_multidevice_gpu.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vbsky_b200.parallel import GLOBAL_KEYS, allreduce_elbo, payload_count, shard_units, unpack_payload

T, K, M, N_TIPS = 7, 3, 5, 6


def pack_reference(unit_tree, ll, grads, T, n, m, K):
    """What vbsky_reduce_elbo computes, in plain torch (test-side reference of the payload layout)."""
    parts = [ll.sum().reshape(1)]
    for k in GLOBAL_KEYS:
        g = grads[k]
        parts.append(g.sum(0).reshape(-1) if g.dim() > 1 else g.sum().reshape(1))
    loc = torch.zeros(T, n - 1, dtype=torch.float64, device=ll.device)
    per_unit = torch.cat([grads["proportions"], grads["root_proportion"].reshape(-1, 1)], 1)
    loc.index_add_(0, unit_tree.long(), per_unit)
    return torch.cat(parts + [loc.reshape(-1)]) / K


def _fake_unit(t, k):
    """Deterministic stand-in for one unit's (ll, gradients)."""
    rng = np.random.default_rng(1000 * t + k)
    return rng.normal(), {"origin": rng.normal(), "clock_rate": rng.normal(), "R": rng.normal(size=M),
                          "delta": rng.normal(size=M), "s": rng.normal(size=M), "rho": rng.normal(size=M),
                          "proportions": rng.normal(size=N_TIPS - 2), "root_proportion": rng.normal()}


def _local(units):
    ll = torch.tensor([_fake_unit(t, k)[0] for t, k in units], dtype=torch.float64)
    grads = {}
    for key in GLOBAL_KEYS + ("proportions", "root_proportion"):
        rows = [np.atleast_1d(_fake_unit(t, k)[1][key]) for t, k in units]
        g = torch.tensor(np.stack(rows), dtype=torch.float64)
        grads[key] = g[:, 0] if key in ("origin", "clock_rate", "root_proportion") else g
    ut = torch.tensor([t for t, _ in units], dtype=torch.int32)
    return ut, ll, grads


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    units = shard_units(T, K, rank, world)
    vec = allreduce_elbo(pack_reference(*_local(units), T, N_TIPS, M, K))
    q.put((rank, len(units), vec.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_units_partition():
    for world in (1, 2, 3, 8):
        seen = sorted(u for r in range(world) for u in shard_units(100, 10, r, world))
        assert seen == [(t, k) for t in range(100) for k in range(10)]
        sizes = [len(shard_units(100, 10, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
        for r in range(world):  # grouped by tree (reduce_elbo's precondition)
            trees = [t for t, _ in shard_units(100, 10, r, world)]
            assert trees == sorted(trees)


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
    want = pack_reference(*_local(shard_units(T, K, 0, 1)), T, N_TIPS, M, K).numpy()
    assert want.shape == (payload_count(T, N_TIPS, M),)
    assert sum(r[1] for r in res) == T * K
    for _, _, got in res:
        np.testing.assert_allclose(got, want, rtol=1e-13, atol=1e-13)
    un = unpack_payload(torch.tensor(want), T, N_TIPS, M)
    assert un["R"].shape == (M,) and un["rho"].shape == (M,) and un["local"].shape == (T, N_TIPS - 1)
    # the local block really needs the exchange: with K = 3 units per tree dealt over 2 ranks, neither rank alone has it
    alone = pack_reference(*_local(shard_units(T, K, 0, 2)), T, N_TIPS, M, K).numpy()
    assert not np.allclose(alone[3 + 4 * M:], want[3 + 4 * M:])


def test_shard_plan_and_tile_ranges():
    """Units are preferred (no in-step exchange) while every rank keeps enough of them; otherwise site-pattern blocks.
    Tile ranges are a contiguous, balanced partition."""
    from vbsky_b200.parallel import shard_plan, tile_range

    assert shard_plan(1000, 117, 1) == (1, 1)
    assert shard_plan(1000, 117, 8) == (8, 1)      # config 3: 125 units per rank
    assert shard_plan(16, 40, 8) == (1, 8)         # config 2: one tree, 16 samples -> 5 tiles per rank
    assert shard_plan(16, 4, 8) == (8, 1)          # fewer tiles than ranks: nothing to split
    assert shard_plan(10000, 8, 8) == (8, 1)       # config 4
    for tiles in (1, 5, 40, 117, 391):
        for parts in (1, 2, 3, 8):
            rs = [tile_range(tiles, p, parts) for p in range(parts)]
            assert rs[0][0] == 0 and rs[-1][1] == tiles
            assert all(a[1] == b[0] for a, b in zip(rs[:-1], rs[1:]))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
