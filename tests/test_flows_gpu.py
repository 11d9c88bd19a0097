"""GPU parity of the local variational family against the torch-fp64 restatement of
prob/transform.py + prob/distribution.py: samples, log q, and gradients of a loss in (x, log q)."""
import numpy as np
import pytest
import torch

import oracle.prob as oprob

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


@pytest.mark.parametrize("T,K,d", [(1, 1, 1), (3, 4, 7), (5, 10, 199)])
def test_local_flow_matches_oracle(cuda, T, K, d):
    from vbsky_b200.flows import local_flow

    rng = np.random.default_rng(T * 100 + d)
    mu = rng.normal(0, 1.5, (T, d))
    ls = rng.normal(-0.5, 0.7, (T, d))
    mu[0, 0], ls[0, 0] = 30.0, -3.0  # saturates the ZeroOne clip (x = 1 - 1e-7)
    eps = rng.standard_normal((T * K, d))
    ut = np.repeat(np.arange(T, dtype=np.int32), K)
    wx = rng.normal(size=(T * K, d))
    wq = rng.normal(size=T * K)

    tmu, tls = torch.tensor(mu, device=cuda, requires_grad=True), torch.tensor(ls, device=cuda, requires_grad=True)
    x, logq = local_flow(tmu, tls, torch.tensor(eps, device=cuda), torch.tensor(ut, device=cuda))
    loss = (x * torch.tensor(wx, device=cuda)).sum() + (logq * torch.tensor(wq, device=cuda)).sum()
    loss.backward()

    omu, ols = torch.tensor(mu, requires_grad=True), torch.tensor(ls, requires_grad=True)
    oloss = 0.0
    for u in range(T * K):
        t = ut[u]
        prm = {"mu": omu[t], "log_sigma": ols[t]}
        ox = oprob.z1_sample(prm, torch.tensor(eps[u]))
        olq = oprob.z1_log_pdf(prm, ox)
        np.testing.assert_allclose(x[u].detach().cpu().numpy(), ox.detach().numpy(), rtol=1e-13, atol=0)
        assert abs(logq[u].item() - olq.item()) <= 1e-9 * abs(olq.item())
        oloss = oloss + (ox * torch.tensor(wx[u])).sum() + olq * wq[u]
    oloss.backward()
    for got, want in ((tmu.grad, omu.grad), (tls.grad, ols.grad)):
        want = want.numpy()
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-6, atol=1e-6 * np.abs(want).max())
    assert x.max().item() <= 1 - 1e-7 and x.min().item() >= 1e-7
