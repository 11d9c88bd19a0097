"""End-to-end run on the reference's HCV example alignment (example/hcv.nexus, 63 sequences x 411 columns; the raw
sequences are committed as a test fixture): subsample -> starting trees on the device -> site patterns -> ensemble
handles -> Adagrad on the batched ELBO.  Mirrors example.ipynb's flow (SeqData.prep_data / setup_flows / loop) with
this package's pieces.  Usage: python examples/hcv_fit.py [n_trees] [n_tips] [n_iter]"""
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

from vbsky_b200.ensemble import Ensemble  # noqa: E402
from vbsky_b200.fit import EnsembleELBO  # noqa: E402
from vbsky_b200.substitution import HKY  # noqa: E402
from vbsky_b200.tips import pad_tips, process_tips  # noqa: E402
from vbsky_b200.treebuild import PackedAlignment, sample_trees  # noqa: E402


M = 10  # skyline intervals (example.ipynb)


def example_flows():
    """The global variational family of example/example.ipynb: R, delta, origin ~ pos, rho_m ~ z1, psi = 0 (all tips are
    sampled at the present with probability rho_m), two GMRF precisions, HCV clock 7.9e-4 substitutions/site/year."""
    return {"origin": ("pos", 1), "origin_start": ("const", 0.0), "delta": ("pos", M), "R": ("pos", M), "rho_m": ("z1", 1),
            "s": ("const", [0.0] * M), "precision_R": ("pos", 1), "precision_delta": ("pos", 1), "clock_rate": ("const", 0.79e-3)}


def example_prior(p):
    """example.ipynb's ``_params_prior_loglik`` on [K, dim] CUDA tensors: gamma hyper-priors on the precisions, Beta(1, 9999)
    on rho_m, log-normal priors on R, delta, origin, GMRF smoothing of log R and log delta."""
    ln = lambda lx, mu_, sig: (math.log(2 * math.pi) + ((lx - mu_) / sig) ** 2) / -2 - lx - math.log(sig)
    gam = lambda x, a, scale: (a - 1) * torch.log(x / scale) - x / scale - math.lgamma(a) - math.log(scale)
    tau = {"R": p["precision_R"][..., 0], "delta": p["precision_delta"][..., 0]}
    ll = gam(tau["R"], 0.001, 1 / 0.001) + gam(tau["delta"], 0.001, 1 / 0.001)
    a, b = 1.0, 9999.0
    x = p["rho_m"]
    ll = ll + ((a - 1) * torch.log(x) + (b - 1) * torch.log1p(-x) - (math.lgamma(a) + math.lgamma(b) - math.lgamma(a + b))).sum(-1)
    for k in ("R", "delta", "origin"):
        ll = ll + ln(torch.log(p[k]), 1.0, 1.25).sum(-1)
    for k in ("R", "delta"):
        lr = torch.log(p[k])
        ll = ll - (tau[k] / 2) * (torch.diff(lr, dim=-1) ** 2).sum(-1)
        ll = ll + (lr.shape[-1] - 1) / 2 * torch.log(tau[k] / (2 * math.pi))
    return ll


def main(n_trees=10, n_tips=10, n_iter=30, seed=0, device="cuda:0", verbose=True):
    with open(os.path.join(ROOT, "tests", "golden", "hcv_patterns.json")) as f:
        fx = json.load(f)
    names, seqs = fx["names"], [s.upper() for s in fx["seqs"]]
    rng = np.random.default_rng(seed)
    dev = torch.device(device)
    sample_times = {s: 0.0 for s in names}  # contemporaneous sample (SeqData(..., contemp=True))
    rows = np.stack([rng.choice(len(names), n_tips, replace=False) for _ in range(n_trees)])  # fasta.py:253
    aln = PackedAlignment(seqs, dev)
    tds, maps, omega = sample_trees(aln, names, sample_times, rows)
    tips = pad_tips([process_tips([names[i] for i in r], [seqs[i] for i in r], maps[t]) for t, r in enumerate(rows)])
    ens = Ensemble(tds, tips, HKY(2.7), device=dev)
    model = EnsembleELBO(ens, K=10, global_flows=example_flows(), prior_fn=example_prior, seed=seed)
    hist = model.fit(n_iter=n_iter, step_size=0.5)
    par = model.params
    R_mean = torch.exp(par["R"]["mu"] + 0.5 * torch.exp(2 * par["R"]["log_sigma"])).detach().cpu().numpy()
    if verbose:
        print(f"{n_trees} trees x {n_tips} tips, {tips[0].patterns.shape[0]} padded site patterns, omega = {omega.round(4).tolist()}")
        print("-ELBO: first %.2f  last %.2f" % (hist[0], hist[-1]))
        print("posterior mean of R per interval:", np.round(R_mean, 3).tolist())
    ens.close()
    return hist, R_mean


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:4]]
    main(*a)
