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
