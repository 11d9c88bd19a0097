"""BASELINE.json's full problem sizes on the GPU, checked through size-independent properties (the oracle
would take minutes there): sum consistency, directional finite differences on the kernel's own value
path, permutation invariance over site patterns, linearity in the pattern counts, bitwise
reproducibility -- plus the oracle on a random subset of patterns at the full tree size."""
import numpy as np
import pytest
import torch

import oracle.prune as oprune
import oracle.substitution as osub
import oracle.tree_data as otd

pytestmark = pytest.mark.gpu

SHAPES = {
    "config2 (500 tips x 10k sites)": (1, 500, 10000, 2, 1.0),
    "config3 (200 tips x 29,903 sites)": (2, 200, 29903, 2, 1.12e-3),
    "config4 (50 tips x 2000 sites, many trees)": (40, 50, 2000, 2, 1.12e-3),
}


@pytest.mark.parametrize("name", list(SHAPES))
def test_fullsize_properties(cuda, name):
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, codes_to_partials
    from vbsky_b200.synth import make_ensemble
    from vbsky_b200.tips import PackedTips

    T, n, S, K, clock = SHAPES[name]
    Q = HKY(2.7)
    data = make_ensemble(T, n, S, K, seed=len(name), Q=Q, clock_rate=clock, device=cuda)
    rng = np.random.default_rng(1)
    counts = [rng.integers(0, 4, S).astype(np.float64) for _ in range(T)]
    tips = [PackedTips(t.patterns, c) for t, c in zip(data.tips, counts)]
    ens = Ensemble(data.tds, tips, Q, device=cuda)
    bl = np.stack([data.heights[t].branch_lengths[k] for t in range(T) for k in range(K)]) * clock
    ut = torch.as_tensor(np.repeat(np.arange(T, dtype=np.int32), K), device=cuda)
    blt = torch.as_tensor(bl, device=cuda)
    ll, dll, site = ens.prune(blt, ut, want_grad=True, want_site_ll=True)
    assert torch.isfinite(ll).all() and torch.isfinite(dll).all() and torch.isfinite(site).all()

    # (1) sum consistency: ll == sum_s counts[s] * site_ll[s]  (bdsky.py:341)
    w = torch.as_tensor(np.stack(counts), device=cuda)[ut.long()]
    np.testing.assert_allclose(ll.cpu().numpy(), (w * site).sum(1).cpu().numpy(), rtol=1e-12)

    # (2) bitwise reproducibility
    ll2, dll2, site2 = ens.prune(blt, ut, want_grad=True, want_site_ll=True)
    assert torch.equal(ll, ll2) and torch.equal(dll, dll2) and torch.equal(site, site2)

    # (3) directional derivative of the kernel's value path vs its analytic gradient (relative step)
    d = torch.as_tensor(rng.uniform(0.5, 1.5, bl.shape), device=cuda) * blt
    eps = 1e-5
    lp, _, _ = ens.prune(blt + eps * d, ut, want_grad=False)
    lm, _, _ = ens.prune(blt - eps * d, ut, want_grad=False)
    fd = ((lp - lm) / (2 * eps)).cpu().numpy()
    an = (dll * d).sum(1).cpu().numpy()
    np.testing.assert_allclose(an, fd, rtol=2e-6, atol=2e-6 * np.abs(an).max())  # central differences: O(eps^2) + rounding

    # (4) the oracle on a random subset of patterns at the full tree size
    u = 0
    td = data.tds[0]
    idx = rng.choice(S, size=48, replace=False)
    part = codes_to_partials(data.tips[0].patterns[idx])
    otdd = otd.TreeData(td.postorder, td.child_parent, td.parent_child, td.sample_times)
    want = oprune.prune_loglik(bl[u], osub.HKY(2.7), part, otdd)
    # absolute floor: the reference forms short-branch P(t) off-diagonals by cancellation
    np.testing.assert_allclose(site[u].cpu().numpy()[idx], want, rtol=1e-9, atol=1e-9)
    ens.close()

    # (5) permutation invariance over patterns and linearity in the counts (fresh ensembles)
    perm = rng.permutation(S)
    tips_p = [PackedTips(t.patterns[perm], 2.0 * c[perm]) for t, c in zip(data.tips, counts)]
    ens_p = Ensemble(data.tds, tips_p, Q, device=cuda)
    llp, dllp, _ = ens_p.prune(blt, ut, want_grad=True)
    np.testing.assert_allclose(llp.cpu().numpy(), 2.0 * ll.cpu().numpy(), rtol=1e-12)
    scale = dll.abs().max().item()
    np.testing.assert_allclose(dllp.cpu().numpy(), 2.0 * dll.cpu().numpy(), rtol=1e-9, atol=1e-11 * scale)
    ens_p.close()


@pytest.mark.parametrize("name", list(SHAPES))
def test_fullsize_complete_units_match_oracle(cuda, name):
    """>= 3 COMPLETE (tree, sample) units at each BASELINE full size: total log-likelihood, every per-site value and the
    full d ll / d bl vector against the C/OpenMP port of prune.py (oracle/prune_ref.c; itself pinned to the reference's
    source on small cases by tests/test_reference_pin.py and to the NumPy restatement by tests/test_oracle_selfcheck.py)."""
    import oracle.cref as cref
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble
    from vbsky_b200.tips import PackedTips

    T, n, S, K, clock = SHAPES[name]
    T = min(T, 3)
    Q = HKY(2.7)
    data = make_ensemble(T, n, S, K, seed=3 + len(name), Q=Q, clock_rate=clock, device=cuda)
    rng = np.random.default_rng(2)
    counts = [rng.integers(0, 4, S).astype(np.float64) for _ in range(T)]
    ens = Ensemble(data.tds, [PackedTips(t.patterns, c) for t, c in zip(data.tips, counts)], Q, device=cuda)
    units = [(t, k) for t in range(T) for k in range(K)][:4] if T > 1 else [(0, 0), (0, 1)]
    if len(units) < 3:  # config 2 has one tree: add a third sample of heights by perturbing the branch lengths
        units.append((0, 0))
    bl = np.stack([data.heights[t].branch_lengths[k] for t, k in units]) * clock
    bl[-1] *= rng.uniform(0.7, 1.3, bl.shape[1])
    ut = torch.as_tensor(np.array([t for t, _ in units], dtype=np.int32), device=cuda)
    ll, dll, site = ens.prune(torch.as_tensor(bl, device=cuda), ut, want_grad=True, want_site_ll=True)
    oQ = osub.HKY(2.7)
    for u, (t, _) in enumerate(units):
        td = data.tds[t]
        otdd = otd.TreeData(td.postorder, td.child_parent, td.parent_child, td.sample_times)
        ll_o, g_o, site_o = cref.prune(otdd, bl[u], oQ, data.tips[t].patterns, counts[t], want_grad=True, want_site_ll=True)
        assert abs(ll[u].item() - ll_o) <= 1e-9 * abs(ll_o), (name, u)
        # absolute floor on per-site values: the reference forms short-branch P(t) off-diagonals by cancellation
        np.testing.assert_allclose(site[u].cpu().numpy(), site_o, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(dll[u].cpu().numpy(), g_o, rtol=1e-6, atol=1e-6 * np.abs(g_o).max())
    ens.close()
