"""Generates the committed fixtures under tests/golden/ from the reference's own example data
(/root/reference/example/hcv.nexus, /root/reference/paper/sims/*/xmls/sim1.xml).  Run in the build
container (where /root/reference is mounted); the GPU box only reads the JSON files.

The reference's JAX code cannot be imported here (no jax / msprime / tskit / Bio wheels, no network), so the
expected values are produced by the oracle (oracle/, a line-by-line CPU restatement of the reference) and
by plain NumPy calls that ARE the reference's statements (np.unique for pattern compression)."""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle.bdsky as obdsky  # noqa: E402
import oracle.prune as oprune  # noqa: E402
import oracle.substitution as osub  # noqa: E402
import oracle.tips as otips  # noqa: E402
import oracle.tree_data as otd  # noqa: E402

REF = "/root/reference"


def hcv():
    names, seqs = otips.read_nexus_matrix(os.path.join(REF, "example", "hcv.nexus"))
    assert len(names) == 63 and all(len(s) == 411 for s in seqs)
    part = np.array([osub.encode_partials(s) for s in seqs]).transpose(1, 0, 2)
    pat, cnt = np.unique(part, axis=0, return_counts=True)  # fasta.py:547
    codes = (pat.astype(np.uint8) << np.arange(4, dtype=np.uint8)).sum(-1)
    json.dump({"source": "example/hcv.nexus", "names": names, "seqs": seqs, "n_patterns": int(len(pat)),
               "counts": cnt.tolist(), "first_column_codes": codes[:, 0].tolist()},
              open(os.path.join(HERE, "hcv_patterns.json"), "w"))
    return names, seqs


def ladder_newick(names, times):
    """Deterministic serially-sampled ladder tree over the given taxa (most recent tip first)."""
    order = np.argsort(times, kind="stable")
    s = names[order[0]]
    h = times[order[0]]
    for i in order[1:]:
        top = max(h, times[i]) + 0.37
        s = f"({s}:{top - h:.6f},{names[i]}:{top - times[i]:.6f})"
        h = top
    return s + ";"


def sims(kind="constant", ntaxa=12, nsites=300):
    xml = open(os.path.join(REF, "paper", "sims", kind, "xmls", "sim1.xml")).read()
    rows = re.findall(r'<sequence[^>]*taxon="([^"]+)"[^>]*value="([^"]+)"', xml)
    assert len(rows) == 100
    rows = rows[:ntaxa]
    names = [r[0] for r in rows]
    seqs = [r[1][:nsites].upper() for r in rows]
    dates = np.array([float(nm.split("_")[1]) for nm in names])
    times = dates.max() - dates
    newick = ladder_newick(names, times)
    td, nm = otd.TreeData.from_newick(newick)
    tip, S = otips.process_tips(names, seqs, nm)
    Q = osub.JC69()
    rng = np.random.default_rng(11)
    n = td.n
    import torch

    torch.set_default_dtype(torch.float64)
    m = 5
    params = dict(
        proportions=torch.tensor(rng.uniform(0.2, 0.8, n - 2)), root_proportion=torch.tensor([0.55]),
        origin=torch.tensor([2.5]), origin_start=torch.tensor([float(td.sample_times.max())]),
        R=torch.tensor(rng.lognormal(0.3, 0.3, m)), delta=torch.full((m,), 1.5, dtype=torch.float64),
        s=torch.full((m,), 0.3, dtype=torch.float64), rho=torch.zeros(m, dtype=torch.float64),
        clock_rate=torch.tensor([0.02], dtype=torch.float64))
    for v in params.values():
        v.requires_grad_(True)
    ll, parts = obdsky.loglik(params, td, tip, Q, return_parts=True)
    ll.backward()
    bl = (parts["branch_lengths"] * params["clock_rate"][0]).detach().numpy()
    site_ll = oprune.prune_loglik(bl, Q, tip.partials, td)
    _, gbl = oprune.prune_loglik_and_grad(bl, Q, tip.partials, td)
    out = {
        "source": f"paper/sims/{kind}/xmls/sim1.xml (first {ntaxa} taxa, first {nsites} sites)", "newick": newick,
        "names": names, "seqs": seqs, "model": "JC69",
        "tree": {"postorder": td.postorder.tolist(), "child_parent": td.child_parent.tolist(),
                 "parent_child": td.parent_child.tolist(), "sample_times": td.sample_times.tolist()},
        "n_patterns": int(S), "counts": tip.counts.tolist(),
        "params": {k: v.detach().numpy().tolist() for k, v in params.items()},
        "grads": {k: v.grad.numpy().tolist() for k, v in params.items() if v.grad is not None},
        "loglik": ll.item(), "tree_loglik": parts["tree"].item(), "data_loglik": parts["data"].item(),
        "node_heights": parts["node_heights"].detach().numpy().tolist(), "branch_lengths_scaled": bl.tolist(),
        "site_ll": site_ll.tolist(), "dll_dbl": (gbl * tip.counts[:, None]).sum(0).tolist(),
    }
    json.dump(out, open(os.path.join(HERE, f"sims_{kind}_sim1.json"), "w"))


if __name__ == "__main__":
    hcv()
    for kind in ("constant", "zigzag"):
        sims(kind)
    print("golden fixtures written to", HERE)
