"""Randomised parity sweep (scripts/stress.py): tree sizes 2..301, pattern counts around the 256-pattern tile
boundary, kappa 0.3..11, up to 30 % ambiguity codes, caterpillars, compressed + padded patterns, zero counts.
Asserted on well-conditioned units (every branch >= 1e-7 substitutions/site); for zero or ~1e-9 branch lengths the
reference's own P @ (exp(t d) * (Pinv @ p)) is rounding noise (clipped 1e-40 vs 1e-17 leftovers) and parity is not defined."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "scripts"))

import oracle.prune as oprune  # noqa: E402
import oracle.substitution as osub  # noqa: E402
import oracle.tree_data as otd  # noqa: E402

pytestmark = pytest.mark.gpu


def test_randomised_sweep_well_conditioned(cuda):
    import stress
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY, codes_to_partials

    checked = 0
    for case in range(60):
        tds, tips, bl, ut, kappa, desc = stress.gen_case(7, case)
        pos = np.where(bl != 0, bl, np.inf).min(axis=1)
        good = [u for u in range(len(ut)) if pos[u] >= 1e-7 and not (bl[u] == 0).any()]
        if not good:
            continue
        ens = Ensemble(tds, tips, HKY(kappa), device=cuda)
        ll, dll, site = ens.prune(torch.as_tensor(bl, device=cuda), torch.as_tensor(ut, device=cuda), want_grad=True, want_site_ll=True)
        ll, dll, site = ll.cpu().numpy(), dll.cpu().numpy(), site.cpu().numpy()
        ens.close()
        oQ = osub.HKY(kappa)
        for u in good[:2]:
            t = ut[u]
            td = otd.TreeData(tds[t].postorder, tds[t].child_parent, tds[t].parent_child, tds[t].sample_times)
            want, g, sll = oprune.data_loglik_and_grad(bl[u], oQ, codes_to_partials(tips[t].patterns), tips[t].counts, td)
            e_v, e_g = stress.errors(ll[u], dll[u], site[u], want, g, sll)
            assert e_v < 1e-9 and e_g < 1e-6, (desc, u, e_v, e_g)
            checked += 1
    assert checked >= 30
