"""GPU parity: CUDA pruning kernel (through the C ABI) vs the NumPy oracle of prune.py.
Tolerances are BASELINE.json's: 1e-9 relative on per-site / total log-likelihoods, 1e-6 relative
on gradients (fp64)."""
import numpy as np
import pytest
import torch

import oracle.prune as oprune
import oracle.substitution as osub
import oracle.tree_data as otd
from oracle.util import TipData

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-9
LL_ATOL = 1e-13  # a fully ambiguous pattern has log-likelihood exactly 0; allow rounding noise there
GRAD_RTOL = 1e-6


def _oracle_td(td):
    return otd.TreeData(td.postorder, td.child_parent, td.parent_child, td.sample_times)


def _check(ens_data, kappa, cuda, K):
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, codes_to_partials

    Q = HKY(kappa)
    oQ = osub.HKY(kappa)
    ens = Ensemble(ens_data.tds, ens_data.tips, Q, device=cuda)
    T = len(ens_data.tds)
    bl = np.stack([ens_data.heights[t].branch_lengths[k] for t in range(T) for k in range(K)]) * ens_data.clock_rate
    unit_tree = np.repeat(np.arange(T, dtype=np.int32), K)
    ll, dll, site = ens.prune(torch.as_tensor(bl, device=cuda), torch.as_tensor(unit_tree, device=cuda),
                              want_grad=True, want_site_ll=True)
    ll_v, _, site_v = ens.prune(torch.as_tensor(bl, device=cuda), torch.as_tensor(unit_tree, device=cuda),
                                want_grad=False, want_site_ll=True)
    torch.cuda.synchronize()
    ll, dll, site = ll.cpu().numpy(), dll.cpu().numpy(), site.cpu().numpy()
    assert np.array_equal(ll, ll_v.cpu().numpy())
    assert np.array_equal(site, site_v.cpu().numpy())
    for u in range(T * K):
        t = unit_tree[u]
        td = _oracle_td(ens_data.tds[t])
        part = codes_to_partials(ens_data.tips[t].patterns)
        tot, g, sll = oprune.data_loglik_and_grad(bl[u], oQ, part, ens_data.tips[t].counts, td)
        np.testing.assert_allclose(site[u], sll, rtol=LL_RTOL, atol=LL_ATOL)
        np.testing.assert_allclose(ll[u], tot, rtol=LL_RTOL, atol=LL_ATOL)
        np.testing.assert_allclose(dll[u], g, rtol=GRAD_RTOL, atol=GRAD_RTOL * np.abs(g).max())
    ens.close()


@pytest.mark.parametrize("T,n,S,K,kappa,clock", [
    (2, 2, 50, 2, 2.7, 0.05),
    (3, 3, 129, 2, 1.0, 0.1),
    (2, 10, 300, 3, 2.7, 0.05),
    (2, 63, 411, 2, 2.7, 1.12e-3),
    (1, 200, 1000, 2, 2.7, 1.12e-3),
    (1, 500, 700, 1, 2.7, 1.0),
])
def test_prune_matches_oracle(cuda, T, n, S, K, kappa, clock):
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    data = make_ensemble(T, n, S, K, seed=n * 1000 + S, Q=HKY(kappa), clock_rate=clock, missing=0.05)
    _check(data, kappa, cuda, K)


def _gtr(rng):
    """A general time-reversible model with non-uniform frequencies as a (pi, P, d, Pinv) tuple: the dense matrix path."""
    pi = rng.dirichlet(np.full(4, 5.0))
    S = np.triu(rng.uniform(0.2, 2.0, (4, 4)), 1)
    S = S + S.T
    Q = S * pi[None, :]
    np.fill_diagonal(Q, -Q.sum(1))
    Q /= -(pi * np.diag(Q)).sum()
    sq = np.sqrt(pi)
    d, V = np.linalg.eigh(sq[:, None] * Q / sq[None, :])  # symmetrised; eigenvectors orthonormal
    return pi, V / sq[:, None], d, V.T * sq[None, :]


@pytest.mark.parametrize("T,n,S,K", [(2, 9, 300, 2), (1, 120, 700, 2)])
def test_prune_general_model_dense_matrices(cuda, T, n, S, K):
    """Non-K80 models take the dense 4x4 path (HKY/JC69 take the structured one): parity with the oracle."""
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, SubstitutionModel, codes_to_partials
    from vbsky_b200.synth import make_ensemble

    rng = np.random.default_rng(n)
    pi, P, d, Pinv = _gtr(rng)
    Q, oQ = SubstitutionModel(pi, P, d, Pinv), osub.SubstitutionModel(pi, P, d, Pinv)
    assert np.allclose(Q.Q.sum(1), 0, atol=1e-12) and np.allclose(pi @ Q.Q, 0, atol=1e-12)
    data = make_ensemble(T, n, S, K, seed=n + S, Q=HKY(2.0), clock_rate=0.05, missing=0.05)
    ens = Ensemble(data.tds, data.tips, Q, device=cuda)
    bl = np.stack([data.heights[t].branch_lengths[k] for t in range(T) for k in range(K)]) * 0.05
    ut = np.repeat(np.arange(T, dtype=np.int32), K)
    ll, dll, site = ens.prune(torch.as_tensor(bl, device=cuda), torch.as_tensor(ut, device=cuda), want_site_ll=True)
    ll, dll, site = ll.cpu().numpy(), dll.cpu().numpy(), site.cpu().numpy()
    ens.close()
    for u in range(T * K):
        t = ut[u]
        tot, g, sll = oprune.data_loglik_and_grad(bl[u], oQ, codes_to_partials(data.tips[t].patterns), data.tips[t].counts, _oracle_td(data.tds[t]))
        np.testing.assert_allclose(site[u], sll, rtol=LL_RTOL, atol=LL_ATOL)
        np.testing.assert_allclose(ll[u], tot, rtol=LL_RTOL)
        np.testing.assert_allclose(dll[u], g, rtol=GRAD_RTOL, atol=GRAD_RTOL * np.abs(g).max())


def test_structured_and_dense_matrix_paths_agree(cuda, monkeypatch):
    """HKY through the K80-structured matrices vs the same model forced through the dense path."""
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    data = make_ensemble(2, 40, 900, 2, seed=4, Q=HKY(2.7), clock_rate=0.02, missing=0.05)
    ens = Ensemble(data.tds, data.tips, HKY(2.7), device=cuda)
    bl = np.stack([data.heights[t].branch_lengths[k] for t in range(2) for k in range(2)]) * 0.02
    bl[1, 3] = 0.0  # one unit through the literal-clip variant
    bl_d = torch.as_tensor(bl, device=cuda)
    ut = torch.as_tensor(np.repeat(np.arange(2, dtype=np.int32), 2), device=cuda)
    a = ens.prune(bl_d, ut, want_site_ll=True)
    monkeypatch.setenv("VBSKY_DENSE_MATRICES", "1")
    b = ens.prune(bl_d, ut, want_site_ll=True)
    monkeypatch.delenv("VBSKY_DENSE_MATRICES")
    ens.close()
    assert not torch.equal(a[1], b[1])  # different arithmetic ...
    for x, y, tol in ((a[0], b[0], 1e-12), (a[2], b[2], 1e-11), (a[1], b[1], 1e-9)):  # ... same numbers
        x, y = x.cpu().numpy(), y.cpu().numpy()
        assert np.max(np.abs(x - y)) <= tol * max(1.0, np.abs(y).max())


def test_prune_compressed_patterns_and_padding(cuda):
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    data = make_ensemble(3, 12, 2000, 2, seed=7, Q=HKY(2.7), clock_rate=0.02, compress=True)
    assert any((t.counts == 0).any() for t in data.tips) or len({t.patterns.shape[0] for t in data.tips}) == 1
    _check(data, 2.7, cuda, 2)


def test_prune_caterpillar_and_ambiguity(cuda):
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, codes_to_partials, encode_codes
    from vbsky_b200.synth import random_tree, sample_heights
    from vbsky_b200.tips import PackedTips

    rng = np.random.default_rng(3)
    n, L = 40, 333
    td = random_tree(n, rng, caterpillar=True)
    sh = sample_heights(td, 2, rng, origin=2.0)
    seqs = ["".join(rng.choice(list("ACGTUNX-?RYMWSKBDHV"), L)) for _ in range(n)]
    codes = encode_codes(seqs)
    tips = PackedTips(codes, rng.integers(0, 5, L).astype(np.float64))
    Q, oQ = HKY(2.7), osub.HKY(2.7)
    ens = Ensemble([td], [tips], Q, device=cuda)
    bl = sh.branch_lengths * 0.3
    ll, dll, site = ens.prune(torch.as_tensor(bl, device=cuda), torch.zeros(2, dtype=torch.int32, device=cuda),
                              want_grad=True, want_site_ll=True)
    for u in range(2):
        tot, g, sll = oprune.data_loglik_and_grad(bl[u], oQ, codes_to_partials(codes), tips.counts, _oracle_td(td))
        np.testing.assert_allclose(site[u].cpu().numpy(), sll, rtol=LL_RTOL, atol=LL_ATOL)
        np.testing.assert_allclose(ll[u].item(), tot, rtol=LL_RTOL)
        np.testing.assert_allclose(dll[u].cpu().numpy(), g, rtol=GRAD_RTOL, atol=GRAD_RTOL * np.abs(g).max())
    ens.close()


def test_prune_zero_length_branches_use_literal_clips(cuda):
    """Branch lengths of exactly 0 (and 1e-30) make P(t) entries vanish, so the 1e-40 clips of
    substitution.py:70 become active; those units take the kernel's literal-clip variant."""
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, codes_to_partials
    from vbsky_b200.synth import evolve_codes, random_tree, sample_heights
    from vbsky_b200.tips import PackedTips

    rng = np.random.default_rng(9)
    n, S = 14, 300
    td = random_tree(n, rng)
    Q, oQ = HKY(2.7), osub.HKY(2.7)
    sh = sample_heights(td, 3, rng, origin=1.0)
    bl = sh.branch_lengths * 0.1
    codes = evolve_codes(td, bl[0], oQ, S, rng, missing=0.02)
    bl[0, [0, 3, n + 1]] = 0.0        # tip and internal branches of zero length
    bl[1, [1, 2]] = 1e-30             # far below the 1e-20 threshold
    # unit 2 stays regular (clip-free variant) in the same launch
    tips = PackedTips(codes, np.ones(S))
    ens = Ensemble([td], [tips], Q, device=cuda)
    ll, dll, site = ens.prune(torch.as_tensor(bl, device=cuda), torch.zeros(3, dtype=torch.int32, device=cuda),
                              want_grad=True, want_site_ll=True)
    for u in range(3):
        tot, g, sll = oprune.data_loglik_and_grad(bl[u], oQ, codes_to_partials(codes), tips.counts, _oracle_td(td))
        assert np.isfinite(sll).all()
        np.testing.assert_allclose(site[u].cpu().numpy(), sll, rtol=LL_RTOL, atol=LL_ATOL)
        np.testing.assert_allclose(ll[u].item(), tot, rtol=LL_RTOL)
        # eq. (9) and its scale-free form agree wherever no clip binds; sites that hit a clip have
        # likelihood ~1e-40 and meaningless derivatives in both, so compare on the clip-free sites
        ok = sll > -80.0
        if ok.all():
            np.testing.assert_allclose(dll[u].cpu().numpy(), g, rtol=GRAD_RTOL, atol=GRAD_RTOL * np.abs(g).max())
    ens.close()


@pytest.mark.parametrize("n,S,caterpillar", [(1500, 384, False), (900, 200, True), (4096, 130, False)])
def test_prune_large_trees_vs_c_port(cuda, n, S, caterpillar):
    """Large / degenerate topologies (config-5 sweep sizes) against the C/OpenMP oracle port, which is fast
    enough there; the NumPy oracle cross-checks the C port in tests/test_oracle_selfcheck.py."""
    import oracle.cref as cref
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import evolve_codes, random_tree, sample_heights
    from vbsky_b200.tips import PackedTips

    rng = np.random.default_rng(n)
    td = random_tree(n, rng, caterpillar=caterpillar)
    Q, oQ = HKY(2.7), osub.HKY(2.7)
    sh = sample_heights(td, 1, rng, origin=1.0)
    bl = sh.branch_lengths * (0.02 if not caterpillar else 0.002)
    codes = evolve_codes(td, bl[0], oQ, S, rng, missing=0.02)
    counts = rng.integers(1, 3, S).astype(np.float64)
    ens = Ensemble([td], [PackedTips(codes, counts)], Q, device=cuda)
    ll, dll, site = ens.prune(torch.as_tensor(bl, device=cuda), torch.zeros(1, dtype=torch.int32, device=cuda),
                              want_grad=True, want_site_ll=True)
    tot, g, sll = cref.prune(_oracle_td(td), bl[0], oQ, codes, counts, True, True)
    np.testing.assert_allclose(site[0].cpu().numpy(), sll, rtol=LL_RTOL, atol=1e-11)
    np.testing.assert_allclose(ll[0].item(), tot, rtol=LL_RTOL)
    np.testing.assert_allclose(dll[0].cpu().numpy(), g, rtol=GRAD_RTOL, atol=GRAD_RTOL * np.abs(g).max())
    ens.close()


@pytest.mark.parametrize("levels,S,units", [(7, 300, 2), (7, 300, 900), (9, 520, 600)])
def test_prune_perfectly_balanced_tree_and_the_deep_stack(cuda, levels, S, units):
    """A perfect binary tree needs log2(n) + 1 parked vectors (8 at 128 tips, 10 at 512): more than the five stack slots
    that fit beside three resident CTAs.  With a couple of waves of work items everything stays in shared memory (fewer
    resident CTAs, fastest walk); with many waves (>= 1776 items) the deepest slots move to the CTA's global deep stack
    and three CTAs per SM stay resident.  Both modes: value, per-site values and the full gradient against the C port."""
    import oracle.cref as cref
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import evolve_codes, sample_heights
    from vbsky_b200.tips import PackedTips
    from vbsky_b200.tree_data import TreeData

    n = 2 ** levels
    edges, nxt = [], 1
    frontier = [("i", 0)]
    for lev in range(levels):
        new = []
        for u in frontier:
            for _ in range(2):
                v = ("t", nxt) if lev == levels - 1 else ("i", nxt)
                nxt += 1
                edges.append((u, v))
                new.append(v)
        frontier = new
    td, _ = TreeData.from_edges(edges)
    assert td.n == n
    rng = np.random.default_rng(levels)
    st = rng.uniform(0.0, 1.0, n)
    st[0] = 0.0
    td = TreeData(td.postorder, td.child_parent, td.parent_child, st)
    Q, oQ = HKY(2.7), osub.HKY(2.7)
    sh = sample_heights(td, 3, rng, origin=1.0)
    bl3 = sh.branch_lengths * 0.03
    bl = bl3[np.arange(units) % 3]  # three distinct samples of heights, repeated
    codes = evolve_codes(td, bl3[0], oQ, S, rng, missing=0.02)
    counts = rng.integers(1, 3, S).astype(np.float64)
    ut = torch.zeros(units, dtype=torch.int32, device=cuda)
    ens = Ensemble([td], [PackedTips(codes, counts)], Q, device=cuda)
    assert max(ens.depth_up, ens.depth_down) > 5
    ll, dll, site = ens.prune(torch.as_tensor(bl, device=cuda), ut, want_grad=True, want_site_ll=True)
    smem = ens.last_stats()["smem"]
    deep_mode = units * ens.tiles >= 4 * 148 * 3
    assert (smem <= 76800) == deep_mode  # three CTAs per SM only in the throughput regime
    for u in range(min(units, 3)):
        tot, g, sll = cref.prune(_oracle_td(td), bl[u], oQ, codes, counts, True, True)
        np.testing.assert_allclose(site[u].cpu().numpy(), sll, rtol=LL_RTOL, atol=1e-11)
        np.testing.assert_allclose(ll[u].item(), tot, rtol=LL_RTOL)
        np.testing.assert_allclose(dll[u].cpu().numpy(), g, rtol=GRAD_RTOL, atol=GRAD_RTOL * np.abs(g).max())
    for u in range(3, units):  # identical inputs give bitwise identical outputs whichever CTA drew the item
        assert torch.equal(ll[u], ll[u % 3]) and torch.equal(dll[u], dll[u % 3])
    ens.close()


def test_prune_many_units_single_pattern(cuda):
    """Ragged extremes: one site pattern (S=1, a single partly-filled tile) and thousands of units."""
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, codes_to_partials
    from vbsky_b200.synth import random_tree, sample_heights
    from vbsky_b200.tips import PackedTips

    rng = np.random.default_rng(11)
    n, U = 9, 3000
    td = random_tree(n, rng)
    Q, oQ = HKY(2.7), osub.HKY(2.7)
    codes = rng.choice(np.array([1, 2, 4, 8, 15], dtype=np.uint8), (1, n))
    ens = Ensemble([td], [PackedTips(codes, np.array([3.0]))], Q, device=cuda)
    bl = sample_heights(td, U, rng, origin=1.0).branch_lengths * 0.3
    ll, dll, site = ens.prune(torch.as_tensor(bl, device=cuda), torch.zeros(U, dtype=torch.int32, device=cuda),
                              want_grad=True, want_site_ll=True)
    for u in (0, 1, U // 2, U - 1):
        tot, g, sll = oprune.data_loglik_and_grad(bl[u], oQ, codes_to_partials(codes), np.array([3.0]), _oracle_td(td))
        np.testing.assert_allclose(ll[u].item(), tot, rtol=LL_RTOL)
        np.testing.assert_allclose(site[u].cpu().numpy(), sll, rtol=LL_RTOL, atol=LL_ATOL)
        np.testing.assert_allclose(dll[u].cpu().numpy(), g, rtol=GRAD_RTOL, atol=GRAD_RTOL * np.abs(g).max())
    ens.close()


@pytest.mark.parametrize("T,n,S,K,clock", [(2, 10, 300, 2, 0.05), (1, 200, 1000, 2, 1.12e-3), (1, 500, 600, 1, 1.0)])
def test_prune_fp32_mode(cuda, T, n, S, K, clock):
    """fp32-storage kernels against the fp64 oracle: BASELINE.json's fp32 tolerance is 1e-5 relative on per-site and
    total log-likelihoods (no gradient tolerance is stated for fp32; 1e-3 relative to the largest entry is checked)."""
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, codes_to_partials
    from vbsky_b200.synth import make_ensemble

    data = make_ensemble(T, n, S, K, seed=n + S, Q=HKY(2.7), clock_rate=clock, missing=0.05)
    ens = Ensemble(data.tds, data.tips, HKY(2.7), device=cuda)
    bl = np.stack([data.heights[t].branch_lengths[k] for t in range(T) for k in range(K)]) * clock
    ut = torch.as_tensor(np.repeat(np.arange(T, dtype=np.int32), K), device=cuda)
    ll, dll, site = ens.prune(torch.as_tensor(bl, device=cuda), ut, want_grad=True, want_site_ll=True, precision="f32")
    for u in range(T * K):
        t = u // K
        tot, g, sll = oprune.data_loglik_and_grad(bl[u], osub.HKY(2.7), codes_to_partials(data.tips[t].patterns),
                                                  data.tips[t].counts, _oracle_td(data.tds[t]))
        np.testing.assert_allclose(site[u].cpu().numpy(), sll, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ll[u].item(), tot, rtol=1e-5)
        np.testing.assert_allclose(dll[u].cpu().numpy(), g, rtol=1e-3, atol=1e-3 * np.abs(g).max())
    ens.close()


def test_prune_loglik_reference_signature(cuda):
    """prune.prune_loglik(branch_lengths, Q, tip_partials, td) with the reference's signature (prune.py:99-131),
    vmapped over patterns by passing [S, n, 4]; its gradient is eq. (9) pulled back through the kernel."""
    from vbsky_b200.prune import prune_loglik
    from vbsky_b200.substitution import HKY, codes_to_partials
    from vbsky_b200.synth import evolve_codes, random_tree, sample_heights

    rng = np.random.default_rng(21)
    n, S = 11, 70
    td = random_tree(n, rng)
    Q, oQ = HKY(2.7), osub.HKY(2.7)
    bl = sample_heights(td, 1, rng, origin=1.0).branch_lengths[0] * 0.2
    part = codes_to_partials(evolve_codes(td, bl, oQ, S, rng, missing=0.05))
    w = rng.normal(size=S)
    tbl = torch.tensor(bl, device=cuda, requires_grad=True)
    ll = prune_loglik(tbl, Q, part, td)
    (ll * torch.tensor(w, device=cuda)).sum().backward()
    want_ll, want_g = oprune.prune_loglik_and_grad(bl, oQ, part, _oracle_td(td))
    np.testing.assert_allclose(ll.detach().cpu().numpy(), want_ll, rtol=LL_RTOL, atol=LL_ATOL)
    g = (want_g * w[:, None]).sum(0)
    np.testing.assert_allclose(tbl.grad.cpu().numpy(), g, rtol=GRAD_RTOL, atol=GRAD_RTOL * np.abs(g).max())
    single = prune_loglik(torch.tensor(bl, device=cuda), Q, part[0], td)  # [n, 4] -> scalar, as in the reference
    assert single.dim() == 0 and abs(single.item() - want_ll[0]) <= LL_RTOL * abs(want_ll[0]) + LL_ATOL


@pytest.mark.parametrize("T,n,S,K,kappa,clock", [(2, 10, 300, 2, 2.7, 0.05), (1, 63, 411, 2, 1.0, 0.01), (1, 200, 1000, 1, 4.0, 1.12e-3)])
def test_kappa_gradient_matches_finite_differences(cuda, T, n, S, K, kappa, clock):
    """d ll / d kappa (north star: gradient w.r.t. substitution parameters; the reference drops these tangents,
    prune.py:149, so the pin is a central difference of the fp64 oracle and of the kernel's own value)."""
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, HKY_dkappa, codes_to_partials
    from vbsky_b200.synth import make_ensemble

    data = make_ensemble(T, n, S, K, seed=n + S, Q=HKY(kappa), clock_rate=clock, missing=0.05)
    bl = np.stack([data.heights[t].branch_lengths[k] for t in range(T) for k in range(K)]) * data.clock_rate
    unit_tree = np.repeat(np.arange(T, dtype=np.int32), K)
    bl_d, ut_d = torch.as_tensor(bl, device=cuda), torch.as_tensor(unit_tree, device=cuda)
    ens = Ensemble(data.tds, data.tips, HKY(kappa), device=cuda)
    # g = d reproduces the branch-length gradient bit for bit
    ll0, dll0, _ = ens.prune(bl_d, ut_d)
    ll1, out1 = ens.prune_spectral_grad(bl_d, ut_d, HKY(kappa).d)
    assert torch.equal(ll0, ll1) and torch.equal(dll0, out1)
    got = ens.dll_dkappa(bl_d, ut_d, kappa).cpu().numpy()
    ens.close()
    h = 1e-4 * kappa
    lls = []
    for kk in (kappa + h, kappa - h):
        e = Ensemble(data.tds, data.tips, HKY(kk), device=cuda)
        lls.append(e.prune(bl_d, ut_d, want_grad=False)[0].cpu().numpy())
        e.close()
    fd_gpu = (lls[0] - lls[1]) / (2 * h)
    scale = np.abs(got).max()
    np.testing.assert_allclose(got, fd_gpu, rtol=1e-6, atol=1e-6 * scale)
    # oracle central difference on the first unit
    td = _oracle_td(data.tds[0])
    part = codes_to_partials(data.tips[0].patterns)
    f = lambda kk: oprune.data_loglik_and_grad(bl[0], osub.HKY(kk), part, data.tips[0].counts, td)[0]
    fd_or = (f(kappa + h) - f(kappa - h)) / (2 * h)
    assert got[0] == pytest.approx(fd_or, rel=1e-6, abs=1e-6 * scale)
    # closed-form spectrum derivative vs a difference of the model's own eigenvalues
    assert np.allclose(HKY_dkappa(kappa), (HKY(kappa + h).d - HKY(kappa - h).d) / (2 * h), rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("T,n,S,K,clock", [(2, 9, 300, 2, 0.05), (1, 40, 700, 2, 0.02)])
def test_gtr_parameter_gradients_match_finite_differences(cuda, T, n, S, K, clock):
    """d ll / d theta for every parameter of a GTR model with non-uniform base frequencies -- the six exchangeabilities
    and the three free frequencies, i.e. directions dQ/dtheta that do NOT commute with Q (north star item 3; unsupported
    in the reference, prune.py:149).  vbsky_prune_param_grad: one kernel evaluation per direction with the per-branch
    matrix D_b = dP_b/dtheta P_b^-1 in eq. (9)'s quadratic form, plus the root-frequency term.  Pinned by central
    differences of the fp64 oracle (NumPy restatement of prune.py) and of the kernel's own value path."""
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import GTR, codes_to_partials, gtr_directions, gtr_rate_matrix
    from vbsky_b200.synth import make_ensemble

    rates = np.array([0.9, 2.6, 0.7, 1.2, 3.1, 1.0])
    pi = np.array([0.31, 0.19, 0.23, 0.27])
    scale = 1.0 / -(pi * np.diag(gtr_rate_matrix(rates, pi))).sum()  # mean rate 1 at the evaluation point
    rates = rates * scale
    Q = GTR(rates, pi)
    data = make_ensemble(T, n, S, K, seed=n + S, Q=Q, clock_rate=clock, missing=0.05)
    bl = np.stack([data.heights[t].branch_lengths[k] for t in range(T) for k in range(K)]) * data.clock_rate
    unit_tree = np.repeat(np.arange(T, dtype=np.int32), K)
    bl_d, ut_d = torch.as_tensor(bl, device=cuda), torch.as_tensor(unit_tree, device=cuda)
    ens = Ensemble(data.tds, data.tips, Q, device=cuda)
    ll0, dll0, _ = ens.prune(bl_d, ut_d)
    # with dQ = Q the per-branch matrix is t_b Q: the branch-length gradient times the branch length
    ll1, tot1, out1, dpi1 = ens.prune_param_grad(bl_d, ut_d, Q.Q)
    assert torch.equal(ll0, ll1)
    np.testing.assert_allclose(out1.cpu().numpy(), (dll0 * bl_d).cpu().numpy(), rtol=1e-9, atol=1e-10 * (dll0 * bl_d).abs().max().item())
    td = _oracle_td(data.tds[0])
    part = codes_to_partials(data.tips[0].patterns)

    def oracle_model(r, p):
        m = GTR(r, p)
        return osub.SubstitutionModel(m.pi, m.P, m.d, m.Pinv)

    def value_gpu(r, p):
        e = Ensemble(data.tds, data.tips, GTR(r, p), device=cuda)
        v = e.prune(bl_d, ut_d, want_grad=False)[0].cpu().numpy()
        e.close()
        return v

    value_or = lambda r, p: oprune.data_loglik_and_grad(bl[0], oracle_model(r, p), part, data.tips[0].counts, td)[0]
    h = 1e-5
    for k, (name, dQ, dpi) in enumerate(gtr_directions(rates, pi)):
        _, got, _, _ = ens.prune_param_grad(bl_d, ut_d, dQ, dpi_dtheta=dpi)
        got = got.cpu().numpy()
        if k < 6:
            rp, rm = rates.copy(), rates.copy()
            rp[k] += h
            rm[k] -= h
            args_p, args_m = (rp, pi), (rm, pi)
        else:
            args_p, args_m = (rates, pi + h * dpi), (rates, pi - h * dpi)
        fd_gpu = (value_gpu(*args_p) - value_gpu(*args_m)) / (2 * h)
        fd_or = (value_or(*args_p) - value_or(*args_m)) / (2 * h)
        sc = max(np.abs(fd_gpu).max(), 1e-8)
        np.testing.assert_allclose(got, fd_gpu, rtol=2e-6, atol=2e-6 * sc, err_msg=name)
        assert got[0] == pytest.approx(fd_or, rel=2e-6, abs=2e-6 * sc), name
    ens.close()
