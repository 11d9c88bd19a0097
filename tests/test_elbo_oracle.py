"""CPU: the oracle's restatement of ``optim.loss`` and of the M-sample ``step`` (oracle/optim.py) against the fixture
produced by the reference's own ``value_and_grad(loss)`` and ``SeqData.loop`` on the HCV example (BASELINE config 1,
tests/golden/config1_hcv_example.json).  Runs without /root/reference, so it also pins the oracle on the GPU box."""
import numpy as np
import torch

import _config1
import oracle.optim as oopt
import oracle.substitution as osub
import oracle.tips as otips
import oracle.tree_data as otd


def _setup():
    g = _config1.load()
    tds, tips = _config1.build_inputs(g, otd.TreeData, otips.process_tips, otips.pad_tips)
    flows = {k: tuple(v) for k, v in g["flows"].items()}
    return g, tds, tips, flows


def _params(glob, loc):
    p = {k: {kk: torch.tensor(v[kk], dtype=torch.float64) for kk in ("mu", "log_sigma")} for k, v in glob.items()}
    p.update({k: {kk: torch.tensor(v[kk], dtype=torch.float64) for kk in ("mu", "log_sigma")} for k, v in loc.items()})
    return p


def test_oracle_step_matches_reference_fixture():
    g, tds, tips, flows = _setup()
    Q = osub.HKY(g["kappa"])
    for j in range(g["n_trees"]):
        want = g["step"]["per_tree"][j]
        p = _params(g["step"]["global_params"], g["step"]["local_params"][j])
        f, grads = oopt.step_value_and_grad(p, flows, tds[j], tips[j], want["eps"], Q, params_prior_loglik=_config1.example_prior_torch)
        assert abs(f - want["loss"]) <= 1e-10 * abs(want["loss"])
        for k, d in want["grad"].items():
            for kk, v in d.items():
                v = np.array(v)
                np.testing.assert_allclose(grads[k][kk].numpy(), v, rtol=1e-7, atol=1e-8 * max(np.abs(v).max(), 1e-300), err_msg=f"{k}.{kk}")


def test_oracle_loop_matches_reference_fixture():
    """Two iterations of SeqData.loop (fasta.py:456-487) restated with the oracle's step + Adagrad rule, replaying the
    recorded base draws: the per-visit losses ``fs`` of the reference's own loop.  Compared over the visits the
    reference itself reproduces to 1e-9 when its draws are perturbed by 1e-13 (``loop.stable_visits``, measured by
    tests/golden/make_golden.py): beyond them the trajectory sits on a discontinuity of the reference's objective (the
    extinction fallback of bdsky.py:62-79) that Adagrad's normalised steps amplify, and no two implementations agree."""
    g, tds, tips, flows = _setup()
    Q = osub.HKY(g["kappa"])
    lp = g["loop"]
    glob = {k: {kk: np.array(v[kk]) for kk in v} for k, v in lp["init_global"].items()}
    loc = [{k: {kk: np.array(v[kk]) for kk in v} for k, v in l.items()} for l in lp["init_local"]]
    leaves = lambda tree: [(k, kk) for k in tree for kk in ("mu", "log_sigma")]
    st_g = {kl: oopt.adagrad_init(glob[kl[0]][kl[1]]) for kl in leaves(glob)}
    st_l = [{kl: oopt.adagrad_init(l[kl[0]][kl[1]]) for kl in leaves(l)} for l in loc]
    assert lp["stable_visits"] >= g["n_trees"] + 3  # at least one full pass over the trees and into the second
    for i in range(lp["n_iter"]):
        for j in range(g["n_trees"]):
            if i * g["n_trees"] + j >= lp["stable_visits"]:
                return
            cur_g = {k: {kk: st_g[(k, kk)][0] for kk in ("mu", "log_sigma")} for k in glob}
            cur_l = {k: {kk: st_l[j][(k, kk)][0] for kk in ("mu", "log_sigma")} for k in loc[j]}
            f, grads = oopt.step_value_and_grad(_params(cur_g, cur_l), flows, tds[j], tips[j], lp["eps"][i][j], Q,
                                                params_prior_loglik=_config1.example_prior_torch)
            want = lp["fs"][j][i]
            assert abs(f - want) <= 1e-9 * abs(want), (i, j, f, want)
            for kl in st_g:
                st_g[kl] = oopt.adagrad_update(grads[kl[0]][kl[1]].numpy(), st_g[kl], lp["step_size"])
            for kl in st_l[j]:
                st_l[j][kl] = oopt.adagrad_update(grads[kl[0]][kl[1]].numpy(), st_l[j][kl], lp["step_size"])
