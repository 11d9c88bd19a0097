"""Site-pattern-block sharding (north star: "shards across trees and site-pattern blocks"): the data term evaluated over
disjoint ranges of 256-pattern tiles gives partial [ll, d/d bl] that add up to the full result, and the split fused
evaluation (forward_partial -> sum of the exchange buffer -> backward_complete) reproduces vbsky_loglik_value_and_grad."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(cuda, T=2, n=24, S=3000, K=2, seed=11):
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    data = make_ensemble(T, n, S, K, seed=seed, Q=HKY(2.7), clock_rate=0.01)
    ens = Ensemble(data.tds, data.tips, HKY(2.7), device=cuda)
    units = [(t, k) for t in range(T) for k in range(K)]
    return data, ens, units


def test_tile_ranges_add_up_to_the_full_data_term(cuda):
    data, ens, units = _setup(cuda)
    assert ens.tiles == 12
    bl = torch.as_tensor(np.stack([data.heights[t].branch_lengths[k] for t, k in units]) * 0.01, device=cuda)
    ut = torch.as_tensor(np.array([t for t, _ in units], dtype=np.int32), device=cuda)
    ll, dll, site = ens.prune(bl, ut, want_grad=True, want_site_ll=True)
    for cuts in ([0, 5, 12], [0, 1, 2, 12], [0, 4, 8, 12]):
        ll_s, dll_s, site_s = torch.zeros_like(ll), torch.zeros_like(dll), torch.zeros_like(site)
        for a, b in zip(cuts[:-1], cuts[1:]):
            l, d, s = ens.prune(bl, ut, want_grad=True, want_site_ll=True, tile_range=(a, b))
            lo, hi = a * 256, min(b * 256, ens.S)
            assert torch.equal(s[:, lo:hi], site[:, lo:hi]) and (s[:, :lo] == 0).all() and (s[:, hi:] == 0).all()
            ll_s += l
            dll_s += d
            site_s += s
        assert torch.equal(site_s, site)
        torch.testing.assert_close(ll_s, ll, rtol=1e-13, atol=0)
        torch.testing.assert_close(dll_s, dll, rtol=1e-11, atol=1e-12 * dll.abs().max().item())
    from vbsky_b200 import _lib

    with pytest.raises(_lib.VbskyError):
        ens.prune(bl, ut, tile_range=(5, 13))
    ens.close()


def _params(data, units, m, cuda, rng):
    U = len(units)
    host = {
        "proportions": np.stack([data.heights[t].proportions[k] for t, k in units]),
        "root_proportion": np.array([data.heights[t].root_proportion[k] for t, k in units]),
        "origin": np.full(U, 0.7), "origin_start": np.full(U, max(td.sample_times.max() for td in data.tds)),
        "clock_rate": np.full(U, 0.01), "R": rng.lognormal(0, 0.3, (U, m)), "delta": np.full((U, m), 3.0),
        "s": np.full((U, m), 0.1), "rho": np.zeros((U, m)),
    }
    return {k: torch.as_tensor(np.ascontiguousarray(v), device=cuda) for k, v in host.items()}


def test_split_fused_evaluation_equals_the_single_call(cuda):
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.parallel import tile_range

    data, ens, units = _setup(cuda)
    m = 5
    params = _params(data, units, m, cuda, np.random.default_rng(0))
    ut = torch.as_tensor(np.array([t for t, _ in units], dtype=np.int32), device=cuda)
    g_full = {k: torch.empty_like(params[k]) for k in Ensemble.GRAD_KEYS}
    full = ens.loglik_device(params, ut, m, grads=g_full)
    full = {k: v.clone() for k, v in full.items()}
    for parts in (2, 3, 12):
        saved = []
        out, g = None, None
        for part in range(parts):
            g = {k: torch.empty_like(params[k]) for k in Ensemble.GRAD_KEYS}

            def exchange(buf, part=part):
                # stand-in for the all-reduce over the site-block group: rank `part` contributes its partial buffer
                saved.append(buf.clone())
                if part == parts - 1:
                    buf.copy_(torch.stack(saved).sum(0))

            out = ens.loglik_device(params, ut, m, grads=g, tile_range=tile_range(ens.tiles, part, parts), exchange=exchange)
        for k in ("ll", "tree", "data"):
            torch.testing.assert_close(out[k], full[k], rtol=1e-13, atol=0)
        for k in Ensemble.GRAD_KEYS:
            torch.testing.assert_close(g[k], g_full[k], rtol=1e-10, atol=1e-11 * max(g_full[k].abs().max().item(), 1e-300))
    ens.close()


def _worker(rank, world, port, q):
    import os

    import torch.distributed as dist

    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.parallel import tile_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    data, ens, units = _setup(dev)
    m = 5
    params = _params(data, units, m, dev, np.random.default_rng(0))
    ut = torch.as_tensor(np.array([t for t, _ in units], dtype=np.int32), device=dev)
    g = {k: torch.empty_like(params[k]) for k in Ensemble.GRAD_KEYS}
    out = ens.loglik_device(params, ut, m, grads=g, tile_range=tile_range(ens.tiles, rank, world), exchange=lambda buf: dist.all_reduce(buf))
    torch.cuda.synchronize()
    q.put((rank, out["ll"].cpu().numpy(), {k: v.cpu().numpy() for k, v in g.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_site_block_sharding_matches_one_gpu(cuda):
    import socket

    import torch.multiprocessing as mp

    from vbsky_b200.ensemble import Ensemble

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    data, ens, units = _setup(cuda)
    m = 5
    params = _params(data, units, m, cuda, np.random.default_rng(0))
    ut = torch.as_tensor(np.array([t for t, _ in units], dtype=np.int32), device=cuda)
    g1 = {k: torch.empty_like(params[k]) for k in Ensemble.GRAD_KEYS}
    want = ens.loglik_device(params, ut, m, grads=g1)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, ll, g in res:
        np.testing.assert_allclose(ll, want["ll"].cpu().numpy(), rtol=1e-13)
        for k in Ensemble.GRAD_KEYS:
            ref = g1[k].cpu().numpy()
            np.testing.assert_allclose(g[k], ref, rtol=1e-10, atol=1e-11 * max(np.abs(ref).max(), 1e-300))
    ens.close()
