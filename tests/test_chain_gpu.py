"""GPU parity of the chain around the pruning kernel: height transform, BDSKY tree density and the
fused bdsky.loglik, values and gradients, against the torch-fp64 oracle (autograd stands in for
JAX autodiff).  Tolerances (BASELINE.json): 1e-9 relative on log-densities, 1e-6 on gradients."""
import json
import os

import numpy as np
import pytest
import torch

import oracle.bdsky as obdsky
import oracle.substitution as osub
import oracle.tree_data as otd
from oracle.util import TipData

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _o(td):
    return otd.TreeData(td.postorder, td.child_parent, td.parent_child, td.sample_times)


def _rand_params(td, m, rng, K, rho_last=0.0):
    n = td.n
    sig = lambda z: 1 / (1 + np.exp(-z))
    out = []
    for _ in range(K):
        rho = np.zeros(m)
        rho[-1] = rho_last
        out.append(dict(
            proportions=sig(rng.standard_normal(n - 2)), root_proportion=np.array([sig(rng.standard_normal())]),
            origin=np.array([rng.uniform(1.0, 3.0)]), origin_start=np.array([float(td.sample_times.max())]),
            clock_rate=np.array([rng.uniform(0.01, 0.1)]), R=rng.lognormal(0.2, 0.4, m),
            delta=rng.uniform(1.0, 3.0, m), s=rng.uniform(0.05, 0.5, m), rho=rho))
    return out


def _oracle_eval(p, td, tip, Q, c, cond=True):
    tp = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
    ll, parts = obdsky.loglik(tp, _o(td), tip, Q, c=c, condition_on_survival=cond, return_parts=True)
    ll.backward()
    g = {k: (v.grad.numpy() if v.grad is not None else np.zeros_like(p[k])) for k, v in tp.items()}
    return ll.item(), {k: v.item() for k, v in parts.items() if k in ("tree", "data")}, g


@pytest.mark.parametrize("n,m,S,rho_last", [(2, 3, 40, 0.0), (6, 4, 64, 0.0), (25, 10, 200, 0.0), (25, 10, 100, 0.3), (120, 50, 150, 0.0)])
def test_loglik_matches_oracle(cuda, n, m, S, rho_last):
    from vbsky_b200 import bdsky
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, codes_to_partials
    from vbsky_b200.synth import evolve_codes, random_tree
    from vbsky_b200.tips import PackedTips, compress_patterns

    rng = np.random.default_rng(n * 7 + m)
    T, K = 2, 2
    Q, oQ = HKY(2.7), osub.HKY(2.7)
    tds, tips, otips_ = [], [], []
    for t in range(T):
        td = random_tree(n, rng)
        if rho_last > 0:  # put a few tips exactly at the present so rho-sampling is exercised
            st = td.sample_times.copy()
            st[rng.choice(n, size=max(1, n // 3), replace=False)] = 0.0
            td = type(td)(td.postorder, td.child_parent, td.parent_child, st)
        codes = evolve_codes(td, np.full(2 * n - 2, 0.05), oQ, S, rng, missing=0.05)
        pat, cnt = compress_patterns(codes)
        tds.append(td), tips.append(PackedTips(pat, cnt.astype(float)))
        otips_.append(TipData(codes_to_partials(pat), cnt.astype(float)))
    from vbsky_b200.tips import pad_tips

    padded = pad_tips(tips)
    ens = Ensemble(tds, padded, Q, device=cuda)
    plist = [(t, p) for t in range(T) for p in _rand_params(tds[t], m, rng, K, rho_last)]
    unit_tree = torch.tensor([t for t, _ in plist], dtype=torch.int32, device=cuda)
    for c in [(False, True, True), (False, True, False), (False, False, True)]:
        params = {k: torch.tensor(np.stack([p[k] for _, p in plist]), device=cuda, requires_grad=(k != "origin_start"))
                  for k in bdsky.PARAM_KEYS if k != "grid"}
        ll, parts = bdsky.loglik(params, ens, unit_tree, c=c, return_parts=True)
        ll.sum().backward()
        for u, (t, p) in enumerate(plist):
            o_ll, o_parts, o_g = _oracle_eval(p, tds[t], otips_[t], oQ, c)
            assert abs(ll[u].item() - o_ll) <= 1e-9 * abs(o_ll), (c, u, ll[u].item(), o_ll)
            if c[1]:
                assert abs(parts["tree"][u].item() - o_parts["tree"]) <= 1e-9 * abs(o_parts["tree"])
            if c[2]:
                assert abs(parts["data"][u].item() - o_parts["data"]) <= 1e-9 * abs(o_parts["data"])
            for k in ("proportions", "root_proportion", "origin", "clock_rate", "R", "delta", "s", "rho"):
                got = params[k].grad[u].cpu().numpy().reshape(-1)
                want = o_g[k].reshape(-1)
                if k == "rho":  # the reference only ever differentiates rho_m (optim.py:28-31)
                    got, want = got[-1:], want[-1:]
                if want.size == 0:
                    continue
                scale = max(np.abs(want).max(), 1e-12)
                np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6 * scale, err_msg=f"{c} unit {u} d/d{k}")
    ens.close()


def test_tree_loglik_signature_and_grads(cuda):
    """_tree_loglik mirror called directly (lam, psi, mu, rho, xs, times) incl. non-equidistant grids."""
    from vbsky_b200 import bdsky
    from vbsky_b200.synth import random_tree, sample_heights

    rng = np.random.default_rng(5)
    n, m, U = 17, 6, 3
    td = random_tree(n, rng)
    sh = sample_heights(td, U, rng, origin=2.0)
    tm = sh.heights.max(1) + 0.5
    xs = tm[:, None] - sh.heights
    grid = np.sort(rng.uniform(0.05, 0.95, (U, m - 1)), axis=1)
    times = np.concatenate([np.zeros((U, 1)), grid * tm[:, None], tm[:, None]], axis=1)
    lam, psi, mu = rng.uniform(1, 4, (U, m)), rng.uniform(0.1, 1, (U, m)), rng.uniform(0.1, 2, (U, m))
    rho = np.zeros((U, m))
    tens = [torch.tensor(a, device=cuda, requires_grad=True) for a in (lam, psi, mu, rho, xs, times)]
    for cond in (True, False):
        for t in tens:
            t.grad = None
        lp = bdsky.tree_loglik(*tens, n=n, condition_on_survival=cond)
        lp.sum().backward()
        for u in range(U):
            ot = [torch.tensor(a[u], requires_grad=True) for a in (lam, psi, mu, rho, xs, times)]
            o = obdsky.tree_loglik(*ot, n, condition_on_survival=cond)
            o.backward()
            assert abs(lp[u].item() - o.item()) <= 1e-9 * abs(o.item())
            for k, (a, b) in enumerate(zip(tens, ot)):
                if k == 3:
                    continue
                want = b.grad.numpy()
                np.testing.assert_allclose(a.grad[u].cpu().numpy(), want, rtol=1e-6, atol=1e-6 * np.abs(want).max())


@pytest.mark.parametrize("kind", ["constant", "zigzag"])
def test_golden_fixture_from_reference_sims(cuda, kind):
    """paper/sims/<kind>/xmls/sim1.xml subset (tests/golden/make_golden.py): tree arrays, compressed patterns,
    loglik parts, per-site log-likelihoods and gradients."""
    from vbsky_b200 import bdsky
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import JC69
    from vbsky_b200.tips import process_tips
    from vbsky_b200.tree_data import TreeData

    g = json.load(open(os.path.join(GOLD, f"sims_{kind}_sim1.json")))
    td, nm = TreeData.from_newick(g["newick"])
    for k in ("postorder", "child_parent", "parent_child", "sample_times"):
        assert np.array_equal(getattr(td, k), np.array(g["tree"][k]))  # bit-exact tree indexing
    tip = process_tips(g["names"], g["seqs"], nm)
    assert tip.patterns.shape[0] == g["n_patterns"] and tip.counts.tolist() == g["counts"]  # bit-exact compression
    ens = Ensemble([td], [tip], JC69(), device=cuda)
    ut = torch.zeros(1, dtype=torch.int32, device=cuda)
    params = {k: torch.tensor(v, device=cuda).reshape(1, -1).requires_grad_(k != "origin_start") for k, v in g["params"].items()}
    ll, parts = bdsky.loglik(params, ens, ut, return_parts=True)
    ll.sum().backward()
    assert abs(ll.item() - g["loglik"]) <= 1e-9 * abs(g["loglik"])
    assert abs(parts["tree"].item() - g["tree_loglik"]) <= 1e-9 * abs(g["tree_loglik"])
    assert abs(parts["data"].item() - g["data_loglik"]) <= 1e-9 * abs(g["data_loglik"])
    for k, v in g["grads"].items():
        if k in ("rho", "origin_start"):
            continue
        want = np.array(v).reshape(-1)
        got = params[k].grad.cpu().numpy().reshape(-1)
        np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6 * max(np.abs(want).max(), 1e-12), err_msg=k)
    bl = torch.tensor(g["branch_lengths_scaled"], device=cuda).reshape(1, -1)
    _, dll, site = ens.prune(bl, ut, want_grad=True, want_site_ll=True)
    np.testing.assert_allclose(site[0].cpu().numpy(), np.array(g["site_ll"]), rtol=1e-9, atol=1e-13)
    want = np.array(g["dll_dbl"])
    np.testing.assert_allclose(dll[0].cpu().numpy(), want, rtol=1e-6, atol=1e-6 * np.abs(want).max())
    ens.close()


@pytest.mark.parametrize("n,m,S", [(6, 4, 64), (25, 10, 120)])
def test_loglik_prespecified_grid_matches_oracle(cuda, n, m, S):
    """equidistant_intervals=False (bdsky.py:317-319): times = tm - grid with times[0] = 0, through the FUSED entry point
    and through the host-buffer entry point; value, every gradient incl. d/d origin (which now flows through
    times = tm - grid) and d/d grid, against the oracle (pinned to the reference in tests/test_reference_pin.py)."""
    from vbsky_b200 import bdsky
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, codes_to_partials
    from vbsky_b200.synth import evolve_codes, random_tree
    from vbsky_b200.tips import PackedTips, compress_patterns, pad_tips

    rng = np.random.default_rng(n + 100 * m)
    T, K = 2, 2
    Q, oQ = HKY(2.7), osub.HKY(2.7)
    tds, tips, otips_ = [], [], []
    for t in range(T):
        td = random_tree(n, rng)
        codes = evolve_codes(td, np.full(2 * n - 2, 0.05), oQ, S, rng, missing=0.05)
        pat, cnt = compress_patterns(codes)
        tds.append(td), tips.append(PackedTips(pat, cnt.astype(float)))
        otips_.append(TipData(codes_to_partials(pat), cnt.astype(float)))
    padded = pad_tips(tips)
    ens = Ensemble(tds, padded, Q, device=cuda)
    plist = []
    for t in range(T):
        for p in _rand_params(tds[t], m, rng, K):
            tm = p["origin"][0] + p["origin_start"][0]
            # a decreasing grid of calendar-style change points: grid[0] is ignored by the reference (times[0] := 0)
            p["grid"] = np.concatenate([[tm * 1.3], np.sort(rng.uniform(0.03 * tm, 0.97 * tm, m - 1))[::-1], [0.0]])
            plist.append((t, p))
    unit_tree = torch.tensor([t for t, _ in plist], dtype=torch.int32, device=cuda)
    keys = [k for k in bdsky.PARAM_KEYS]
    for c in [(False, True, True), (False, True, False)]:
        params = {k: torch.tensor(np.stack([p[k] for _, p in plist]), device=cuda, requires_grad=(k != "origin_start")) for k in keys}
        ll = bdsky.loglik(params, ens, unit_tree, c=c, equidistant_intervals=False)
        ll.sum().backward()
        for u, (t, p) in enumerate(plist):
            tp = {k: torch.tensor(v, requires_grad=True) for k, v in p.items()}
            o = obdsky.loglik(tp, _o(tds[t]), otips_[t], oQ, c=c, equidistant_intervals=False)
            o.backward()
            assert abs(ll[u].item() - o.item()) <= 1e-9 * abs(o.item()), (c, u)
            for k in ("proportions", "root_proportion", "origin", "clock_rate", "R", "delta", "s", "grid"):
                want = tp[k].grad.numpy().reshape(-1) if tp[k].grad is not None else np.zeros(p[k].size)
                got = params[k].grad[u].cpu().numpy().reshape(-1)
                if want.size:
                    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6 * max(np.abs(want).max(), 1e-12), err_msg=f"{c} {u} {k}")
    # the equidistant evaluation of the same parameters differs (the grid is really used) ...
    eq = bdsky.loglik({k: v.detach() for k, v in params.items() if k != "grid"}, ens, unit_tree, c=(False, True, False))
    assert not torch.allclose(eq, ll.detach(), rtol=1e-6)
    # ... and the host-buffer entry point takes the grid as well
    host = {k: np.ascontiguousarray(np.stack([p[k] for _, p in plist]).reshape(len(plist), -1)) for k in keys}
    for k in ("root_proportion", "origin", "origin_start", "clock_rate"):
        host[k] = host[k].reshape(-1)
    hg = {k: np.empty_like(host[k]) for k in Ensemble.GRAD_KEYS + ("grid",)}
    out = ens.loglik_host(host, unit_tree.cpu().numpy(), m, flags=1 | 4, grads=hg)
    np.testing.assert_allclose(out["ll"], ll.detach().cpu().numpy(), rtol=1e-12)
    np.testing.assert_allclose(hg["grid"], params["grid"].grad.cpu().numpy(), rtol=1e-10, atol=1e-12)
    ens.close()
