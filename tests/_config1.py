"""Shared loader for tests/golden/config1_hcv_example.json: BASELINE config 1 (example/hcv.nexus set up as
example/example.ipynb) with expected values produced by the reference's own source (tests/golden/make_golden.py)."""
import json
import math
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOCAL = ("proportions", "root_proportion")


def load():
    g = json.load(open(os.path.join(GOLD, "config1_hcv_example.json")))
    h = json.load(open(os.path.join(GOLD, "hcv_patterns.json")))
    g["names"], g["seqs"] = h["names"], [s.upper() for s in h["seqs"]]
    return g


def build_inputs(g, TreeData, process_tips, pad_tips):
    """Trees (from the stored newicks, through ``TreeData.from_newick``) and padded tip data for either the oracle's or the
    product's classes; asserts the tree arrays and pattern counts equal the reference's bit for bit."""
    tds, tips = [], []
    for t, rows in enumerate(g["rows"]):
        td, nm = TreeData.from_newick(g["newicks"][t])
        td.sample_times[:] = 0.0  # contemporaneous sample (fasta.py:529-537 with dates = 1993.0)
        for k in ("postorder", "child_parent", "parent_child", "sample_times"):
            assert np.array_equal(getattr(td, k), np.array(g["trees"][t][k])), (t, k)
        names, seqs = [g["names"][i] for i in rows], [g["seqs"][i] for i in rows]
        out = process_tips(names, seqs, nm)
        tips.append(out[0] if isinstance(out, tuple) and not hasattr(out, "_fields") else out)
        tds.append(td)
    tips = pad_tips(tips)
    for t, tp in enumerate(tips):
        assert np.asarray(tp.counts).tolist() == g["counts"][t]
    return tds, tips


def example_prior_torch(p):
    """example.ipynb's ``_params_prior_loglik`` on torch tensors with an optional leading sample axis."""
    import torch

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
