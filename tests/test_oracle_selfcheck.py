"""Validation of the oracle itself (CPU only).  The reference holds no golden vector for prune_loglik,
its gradient, _tree_loglik or loglik ("parity unpinned"), so the restatement is checked
against independent computations: brute-force enumeration, the closed-form JC69 two-tip likelihood,
central finite differences, the C/OpenMP port and the committed golden fixtures."""
import json
import os

import numpy as np
import pytest
import torch

torch.set_default_dtype(torch.float64)

import oracle.bdsky as obdsky  # noqa: E402
import oracle.cref as cref  # noqa: E402
import oracle.prune as oprune
import oracle.substitution as osub
import oracle.tree_data as otd
from oracle.util import TipData

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _tree():
    return otd.TreeData.from_newick("(((a:1,b:2):1,c:1.5):0.5,(d:1,(e:1,f:1):0.3):1);")[0]


def test_prune_matches_brute_force():
    rng = np.random.default_rng(0)
    td = _tree()
    for kappa, seq in ((2.7, "ACGTNR"), (1.0, "AAAAAA"), (0.5, "YKT-CG")):
        Q = osub.HKY(kappa)
        bl = rng.uniform(0.05, 0.6, 2 * td.n - 2)
        tips = osub.encode_partials(seq)
        assert abs(oprune.prune_loglik(bl, Q, tips, td) - oprune.brute_force_loglik(bl, Q, tips, td)) < 1e-13


def test_jc69_two_tip_closed_form():
    td = otd.TreeData.from_newick("(a:1,b:1)")[0]
    Q = osub.JC69()
    for t1, t2 in ((0.1, 0.3), (1e-4, 2e-4), (2.0, 0.5)):
        d = t1 + t2
        rate = 4.0 / 3.0  # JC69 as normalised by substitution.py: off-diagonal 1/3, diagonal -1
        same = 0.25 * (0.25 + 0.75 * np.exp(-rate * d))
        diff = 0.25 * (-0.25 * np.expm1(-rate * d))
        bl = np.array([t1, t2])
        assert abs(oprune.prune_loglik(bl, Q, osub.encode_partials("AA"), td) - np.log(same)) < 1e-13
        # the reference forms P(t) off-diagonals as a difference of O(1) terms: ~1e-16/t relative noise
        assert abs(oprune.prune_loglik(bl, Q, osub.encode_partials("AG"), td) - np.log(diff)) < max(1e-12, 2e-15 / d)


def test_gradient_matches_finite_differences():
    rng = np.random.default_rng(1)
    td = _tree()
    Q = osub.HKY(2.7)
    bl = rng.uniform(0.05, 0.6, 2 * td.n - 2)
    tips = osub.encode_partials("ACGTNR")
    _, g = oprune.prune_loglik_and_grad(bl, Q, tips, td)
    eps = 1e-6
    for i in range(len(bl)):
        e = np.zeros_like(bl)
        e[i] = eps
        fd = (oprune.prune_loglik(bl + e, Q, tips, td) - oprune.prune_loglik(bl - e, Q, tips, td)) / (2 * eps)
        assert abs(fd - g[i]) < 1e-8


def test_c_port_matches_numpy_oracle():
    from vbsky_b200.synth import evolve_codes, random_tree, sample_heights
    from vbsky_b200.substitution import codes_to_partials

    rng = np.random.default_rng(2)
    Q = osub.HKY(2.7)
    for n in (2, 9, 50):
        td0 = random_tree(n, rng)
        td = otd.TreeData(td0.postorder, td0.child_parent, td0.parent_child, td0.sample_times)
        bl = sample_heights(td0, 1, rng, origin=1.0).branch_lengths[0] * 0.1
        codes = evolve_codes(td0, bl, Q, 257, rng, missing=0.05)
        counts = rng.integers(0, 4, 257).astype(float)
        ll, g, site = cref.prune(td, bl, Q, codes, counts, True, True)
        tot, g2, s2 = oprune.data_loglik_and_grad(bl, Q, codes_to_partials(codes), counts, td)
        np.testing.assert_allclose(site, s2, rtol=1e-12, atol=1e-13)
        assert abs(ll - tot) <= 1e-12 * abs(tot)
        np.testing.assert_allclose(g, g2, rtol=1e-9, atol=1e-10 * np.abs(g2).max())


@pytest.mark.parametrize("kind", ["constant", "zigzag"])
def test_oracle_reproduces_golden_fixture(kind):
    g = json.load(open(os.path.join(GOLD, f"sims_{kind}_sim1.json")))
    td, nm = otd.TreeData.from_newick(g["newick"])
    for k in ("postorder", "child_parent", "parent_child", "sample_times"):
        assert np.array_equal(getattr(td, k), np.array(g["tree"][k]))
    import oracle.tips as otips

    tip, S = otips.process_tips(g["names"], g["seqs"], nm)
    assert S == g["n_patterns"] and tip.counts.tolist() == g["counts"]
    params = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in g["params"].items()}
    ll, parts = obdsky.loglik(params, td, tip, osub.JC69(), return_parts=True)
    assert abs(ll.item() - g["loglik"]) <= 1e-12 * abs(g["loglik"])
    assert abs(parts["tree"].item() - g["tree_loglik"]) <= 1e-12 * abs(g["tree_loglik"])
    ll.backward()
    for k, v in g["grads"].items():
        np.testing.assert_allclose(params[k].grad.numpy(), np.array(v), rtol=1e-10, atol=1e-12)


def test_bdsky_oracle_gradients_match_finite_differences():
    rng = np.random.default_rng(3)
    td = _tree()
    n, m = td.n, 4
    seqs = ["".join(rng.choice(list("ACGT"), 30)) for _ in range(n)]
    part = np.array([osub.encode_partials(s) for s in seqs]).transpose(1, 0, 2)
    pat, cnt = np.unique(part, axis=0, return_counts=True)
    tip = TipData(pat, cnt.astype(float))
    p = dict(proportions=torch.tensor(rng.uniform(0.2, 0.8, n - 2)), root_proportion=torch.tensor([0.6]),
             origin=torch.tensor([5.0]), origin_start=torch.tensor([float(td.sample_times.max()) + 0.5]),
             R=torch.tensor(rng.lognormal(0, 0.5, m)), delta=torch.full((m,), 2.0, dtype=torch.float64),
             s=torch.full((m,), 0.1, dtype=torch.float64), rho=torch.zeros(m, dtype=torch.float64),
             clock_rate=torch.tensor([0.05], dtype=torch.float64))
    for v in p.values():
        v.requires_grad_(True)
    Q = osub.HKY(2.7)
    ll = obdsky.loglik(p, td, tip, Q)
    ll.backward()
    eps = 1e-6
    for k in ("proportions", "root_proportion", "origin", "R", "delta", "s", "clock_rate"):
        for i in range(p[k].numel()):
            with torch.no_grad():
                o = p[k][i].item()
                p[k][i] = o + eps
                lp = obdsky.loglik(p, td, tip, Q).item()
                p[k][i] = o - eps
                lm = obdsky.loglik(p, td, tip, Q).item()
                p[k][i] = o
            fd = (lp - lm) / (2 * eps)
            assert abs(fd - p[k].grad[i].item()) <= 1e-6 * max(1.0, abs(fd)), (k, i)


def test_local_flow_oracle_change_of_variables():
    """prob/transform.py restatement: away from the ZeroOne clip, log q(x) = log N(eps) - sum(log_sigma) -
    sum(log x + log(1-x)) (density of expit(mu + sigma*eps) by the change-of-variables formula)."""
    import oracle.prob as oprob

    rng = np.random.default_rng(4)
    for d in (1, 5, 40):
        prm = {"mu": torch.tensor(rng.normal(0, 1, d)), "log_sigma": torch.tensor(rng.normal(-0.3, 0.5, d))}
        eps = torch.tensor(rng.standard_normal(d))
        x = oprob.z1_sample(prm, eps)
        want = (-0.5 * eps**2 - 0.5 * np.log(2 * np.pi)).sum() - prm["log_sigma"].sum() - (torch.log(x) + torch.log(1 - x)).sum()
        assert abs(oprob.z1_log_pdf(prm, x).item() - want.item()) <= 1e-12 * abs(want.item())
        # round trip through the inverse map
        z = oprob.zero_one_inverse(x)
        np.testing.assert_allclose(oprob.diagonal_affine_inverse(prm, z).numpy(), eps.numpy(), rtol=1e-9, atol=1e-12)
