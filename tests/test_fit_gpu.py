"""End-to-end: the batched ELBO (local flows + heights + BDSKY prior + pruning, all through the CUDA path) is
differentiable w.r.t. every variational parameter and an Adagrad loop improves it (mirror of SeqData.loop)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_fit_improves_elbo(cuda):
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.fit import EnsembleELBO
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    data = make_ensemble(4, 12, 400, 1, seed=5, Q=HKY(2.7), clock_rate=0.05, origin=1.0, compress=True)
    ens = Ensemble(data.tds, data.tips, HKY(2.7), device=cuda)
    model = EnsembleELBO(ens, K=8, m=5, delta=2.0, s=0.2, clock_rate=0.05, seed=1)
    first = np.mean([-model.elbo().item() for _ in range(4)])
    hist = model.fit(n_iter=40, step_size=0.1)
    assert np.isfinite(hist).all()
    last = np.mean([-model.elbo().item() for _ in range(4)])
    assert last < first - 1.0, (first, last)
    for v in model.leaves():
        assert v.grad is not None and torch.isfinite(v.grad).all()
    ens.close()


def test_alignment_to_fit_pipeline(cuda):
    """prep_data end to end on the device (fasta.py:333-360): subsample -> starting trees -> tip patterns ->
    ensemble handles -> a few Adagrad steps of the batched ELBO."""
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.fit import EnsembleELBO
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import evolve_codes, random_tree, sample_heights
    from vbsky_b200.tips import pad_tips, process_tips
    from vbsky_b200.treebuild import PackedAlignment, sample_trees

    rng = np.random.default_rng(2)
    N, L, T, n = 30, 600, 3, 10
    td = random_tree(N, rng)
    bl = sample_heights(td, 1, rng, origin=1.0).branch_lengths[0] * 0.05
    codes = evolve_codes(td, bl, HKY(2.7), L, rng)  # [L, N] masks
    letters = np.full(16, "N")
    letters[[1, 2, 4, 8]] = list("ACGT")
    seqs = ["".join(letters[codes[:, i]]) for i in range(N)]
    names = [f"s{i}" for i in range(N)]
    sample_times = {names[i]: float(td.sample_times[i]) for i in range(N)}
    rows = np.stack([rng.choice(N, n, replace=False) for _ in range(T)])
    aln = PackedAlignment(seqs, cuda)
    tds, maps, omega = sample_trees(aln, names, sample_times, rows)
    assert np.isfinite(omega).all()
    tips = pad_tips([process_tips([names[i] for i in rows[t]], [seqs[i] for i in rows[t]], maps[t]) for t in range(T)])
    ens = Ensemble(tds, tips, HKY(2.7), device=cuda)
    model = EnsembleELBO(ens, K=4, m=5, delta=2.0, s=0.2, clock_rate=0.05, seed=3)
    hist = model.fit(n_iter=5, step_size=0.1)
    assert np.isfinite(hist).all()
    ens.close()


def test_adagrad_kernel_matches_jax_rule(cuda):
    """vbsky_adagrad_update vs the restated jax.example_libraries.optimizers.adagrad + nan_to_num (fasta.py:388,421-433)."""
    import oracle.optim as ooptim
    from vbsky_b200.optim import Adagrad

    rng = np.random.default_rng(0)
    n = 10007
    x0 = rng.standard_normal(n)
    x = torch.as_tensor(x0.copy(), device=cuda)
    opt = Adagrad(x, step_size=0.7)
    state = ooptim.adagrad_init(x0)
    for it in range(6):
        g = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)
        g[rng.random(n) < 0.01] = np.nan
        g[rng.random(n) < 0.01] = np.inf
        g[rng.random(n) < 0.01] = -np.inf
        if it < 2:
            g[:100] = 0.0  # g_sq stays 0 -> no update there
        opt.update(torch.as_tensor(g, device=cuda))
        state = ooptim.adagrad_update(g, state, 0.7)
    assert np.array_equal(x.cpu().numpy(), state[0])
    assert np.array_equal(opt.g_sq.cpu().numpy(), state[1])
    assert np.array_equal(opt.m.cpu().numpy(), state[2])
    assert np.array_equal(x.cpu().numpy()[:100] != x0[:100], np.ones(100, bool))  # moved once gradients appeared


def test_sequential_fit_follows_reference_schedule(cuda):
    """fit(sequential=True): trees visited one at a time; a visit touches the global block and that tree's locals only."""
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.fit import EnsembleELBO
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    T, n, m = 3, 8, 4
    data = make_ensemble(T, n, 300, 1, seed=9, Q=HKY(2.7), clock_rate=0.05, origin=1.0, compress=True)
    ens = Ensemble(data.tds, data.tips, HKY(2.7), device=cuda)
    model = EnsembleELBO(ens, K=6, m=m, delta=2.0, s=0.2, clock_rate=0.05, seed=2)
    G, d = 2 * (1 + m), n - 1
    model.theta.grad = None
    (-model.elbo(trees=[1])).backward()
    g = model.theta.grad.clone()
    mu, ls = g[G : G + T * d].view(T, d), g[G + T * d :].view(T, d)
    assert (mu[1] != 0).any() and (ls[1] != 0).any() and (g[:G] != 0).all()
    assert (mu[[0, 2]] == 0).all() and (ls[[0, 2]] == 0).all()
    first = np.mean([[-model.elbo(trees=[j]).item() for j in range(T)] for _ in range(4)], axis=0)
    hist = np.array(model.fit(n_iter=25, step_size=0.1, sequential=True))
    assert hist.shape == (25, T) and np.isfinite(hist).all()
    last = np.mean([[-model.elbo(trees=[j]).item() for j in range(T)] for _ in range(4)], axis=0)
    assert (last < first).all(), (first, last)
    ens.close()


def test_hcv_example_runs(cuda):
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("hcv_fit", os.path.join(os.path.dirname(__file__), "..", "examples", "hcv_fit.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    hist, R_mean = mod.main(n_trees=4, n_tips=10, n_iter=12, verbose=False)
    assert np.isfinite(hist).all() and np.isfinite(R_mean).all() and np.mean(hist[-3:]) < np.mean(hist[:3])
