"""GPU: ELBO-level parity on BASELINE config 1 -- example/hcv.nexus set up as example/example.ipynb
(variational delta / R / rho_m / origin, psi = 0, rho-sampling at the present, HKY(2.7), clock 7.9e-4).

Expected values come from the reference's own ``value_and_grad(loss)`` (optim.py:35-78, fasta.py:74-78) and from two
iterations of its own ``SeqData.loop`` (fasta.py:387-487), executed under oracle/refshim and committed as
tests/golden/config1_hcv_example.json; the recorded base draws are replayed through ``EnsembleELBO``.
Tolerances (north star): 1e-9 relative on the ELBO, 1e-6 relative on gradients."""
import numpy as np
import pytest
import torch

import _config1

pytestmark = pytest.mark.gpu


def _model(cuda, g):
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.fit import EnsembleELBO
    from vbsky_b200.substitution import HKY
    from vbsky_b200.tips import pad_tips, process_tips
    from vbsky_b200.tree_data import TreeData

    tds, tips = _config1.build_inputs(g, TreeData, process_tips, pad_tips)
    assert tips[0].patterns.shape[0] == g["n_patterns_padded"]
    ens = Ensemble(tds, tips, HKY(g["kappa"]), device=cuda)
    flows = {k: tuple(v) for k, v in g["flows"].items() if k not in _config1.LOCAL}
    return ens, EnsembleELBO(ens, K=10, global_flows=flows, prior_fn=_config1.example_prior_torch)


def _eps(samples, n, cuda):
    """M recorded single-sample draws -> {flow: [M, dim], "local": [M, n-1]} on the device."""
    out = {k: torch.tensor([s[k] for s in samples], device=cuda) for k in samples[0] if k not in _config1.LOCAL}
    out["local"] = torch.tensor([s["proportions"] + s["root_proportion"] for s in samples], device=cuda)
    return out


def test_elbo_and_gradient_match_reference_step(cuda):
    g = _config1.load()
    ens, model = _model(cuda, g)
    n, T = g["n_tips"], g["n_trees"]
    model.set_params(g["step"]["global_params"], g["step"]["local_params"])
    par = model.params
    for j in range(T):
        want = g["step"]["per_tree"][j]
        model.theta.grad = None
        loss = -model.elbo(trees=[j], eps=_eps(want["eps"], n, cuda))
        loss.backward()
        assert abs(loss.item() - want["loss"]) <= 1e-9 * abs(want["loss"]), (j, loss.item(), want["loss"])
        # views of theta.grad in the same layout as the parameters
        gr, off = {}, 0
        for name, sh in model._shapes.items():
            gr[name] = {}
            for k in ("mu", "log_sigma"):
                cnt = int(np.prod(sh))
                gr[name][k] = model.theta.grad[off: off + cnt].view(sh).cpu().numpy()
                off += cnt
        for name, d in want["grad"].items():
            for k, v in d.items():
                v = np.array(v)
                if name == "proportions":
                    got = gr["local"][k][j, : n - 2]
                elif name == "root_proportion":
                    got = gr["local"][k][j, n - 2:]
                else:
                    got = gr[name][k]
                np.testing.assert_allclose(got, v, rtol=1e-6, atol=1e-6 * max(np.abs(v).max(), 1e-300), err_msg=f"tree {j} {name}.{k}")
        others = [t for t in range(T) if t != j]
        assert (gr["local"]["mu"][others] == 0).all()  # a visit touches that tree's local parameters only
    del par
    ens.close()


def test_sequential_fit_reproduces_reference_loop(cuda):
    """The per-visit losses ``fs`` of the reference's own SeqData.loop, replayed with the same initial parameters and
    base draws through EnsembleELBO.fit(sequential=True) (kernels + fused Adagrad)."""
    g = _config1.load()
    ens, model = _model(cuda, g)
    lp = g["loop"]
    model.set_params(lp["init_global"], lp["init_local"])
    hist = model.fit(n_iter=lp["n_iter"], step_size=lp["step_size"], sequential=True,
                     eps_fn=lambda i, j: _eps(lp["eps"][i][j], g["n_tips"], cuda))
    want = np.array(lp["fs"]).T.reshape(-1)  # fs[j][i] -> visits in iteration-major order
    got = np.array(hist).reshape(-1)
    # the visits the reference itself reproduces to 1e-9 under a 1e-13 perturbation of its draws (loop.stable_visits): beyond
    # them the trajectory sits on a discontinuity of the reference's objective (bdsky.py:62-79) and nothing is comparable
    nv = lp["stable_visits"]
    assert nv >= g["n_trees"] + 3
    np.testing.assert_allclose(got[:nv], want[:nv], rtol=1e-9)
    assert np.isfinite(got).all()
    ens.close()


def test_batched_elbo_is_the_sum_of_per_tree_terms(cuda):
    """elbo() over all trees at common global draws = sum_j elbo(trees=[j]) minus the (T-1) extra copies of the global
    entropy and prior terms."""
    g = _config1.load()
    ens, model = _model(cuda, g)
    T, n, K = g["n_trees"], g["n_tips"], model.K
    model.set_params(g["step"]["global_params"], g["step"]["local_params"])
    eps = model.sample_eps(T)
    whole = model.elbo(eps=eps).item()
    parts = 0.0
    for j in range(T):
        e = {k: v for k, v in eps.items() if k != "local"}
        e["local"] = eps["local"][j * K: (j + 1) * K]
        parts += model.elbo(trees=[j], eps=e).item()
    # global-only part: entropy + prior, evaluated by switching the likelihood off through a zero-tree call is not
    # possible; recompute it directly
    from vbsky_b200.fit import _flow_sample

    par = model.params
    glob = 0.0
    samples = {}
    for name, (kind, _) in model.flows.items():
        if kind == "const":
            samples[name] = model._const[name].reshape(1, -1).expand(K, -1)
            continue
        samples[name], lq = _flow_sample(kind, par[name], eps[name])
        glob -= lq.mean().item()
    glob += _config1.example_prior_torch(samples).mean().item()
    assert abs(whole - (parts - (T - 1) * glob)) <= 1e-10 * abs(whole)
    ens.close()
