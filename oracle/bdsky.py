"""Oracle restatement of /root/reference/vbsky/bdsky.py (test infrastructure).

Written against PyTorch fp64 tensors so that ``torch.autograd`` stands in for JAX autodiff
(``where`` has the same evaluate-both-branches semantics in both).  Follows bdsky.py:18-216
(``_tree_loglik``), :219-231, :268-345 (``loglik``), and tree_data.py:287-367 for the height
transform inside the differentiable chain.

Out-of-bounds gathers: ``jnp.take`` is restated with *clip* semantics: the most
recent leaf has searchsorted index m+1, and ``take(A, m)``, ``take(times, m+1)``, ``take(psi, m)``
must clamp for the reference's own example to be finite.
"""
import math

import numpy as np
import torch

from . import prune as oprune
from .util import order_events

F64 = torch.float64


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=F64)


def _take(x, i):
    return x[torch.clamp(i, 0, x.shape[0] - 1)]


def _logit(p):
    return torch.log(p / (1 - p))  # jax.scipy.special.logit


def _isclose(a, b, rtol=1e-5, atol=1e-8):
    return (a - b).abs() <= atol + rtol * b.abs()


def _xlogy(x, y):
    x = x.to(F64)
    return torch.where(x == 0, torch.zeros_like(y * x), x * torch.log(torch.where(x == 0, torch.ones_like(y), y)))


def _xlog1py(x, y):
    x = x.to(F64)
    return torch.where(x == 0, torch.zeros_like(y * x), x * torch.log1p(torch.where(x == 0, torch.zeros_like(y), y)))


def tree_loglik(lam, psi, mu, rho, xs, times, n, condition_on_survival=True, return_parts=False):
    """bdsky.py:18-216 (Stadler et al. 2013, Thm 1).  ``n`` = number of tips (td.n)."""
    lam, psi, mu, rho, xs, times = map(_t, (lam, psi, mu, rho, xs, times))
    m = len(lam)
    assert m == len(psi) == len(mu)
    assert m + 1 == len(times)

    dt = times[1:] - times[:-1]
    A = torch.sqrt((lam - mu - psi) ** 2 + (4 * lam * psi))

    # bdsky.py:48-120 reverse scan
    pi1 = torch.tensor(1.0, dtype=F64)
    logit_pi1 = torch.tensor(0.0, dtype=F64)
    B = [None] * m
    for i in range(m - 1, -1, -1):
        B_i = ((1 - 2 * (1 - rho[i]) * pi1) * lam[i] + mu[i] + psi[i]) / A[i]
        Adt = A[i] * dt[i]
        f = ((1 + B_i) - torch.exp(-Adt) * (1 - B_i)) / ((1 + B_i) + torch.exp(-Adt) * (1 - B_i))
        g = f
        p_i = torch.minimum((lam[i] + mu[i] + psi[i] - A[i] * g) / (2 * lam[i]), torch.tensor(1.0, dtype=F64))
        q = dt[i] * torch.minimum(lam[i] - mu[i], torch.tensor(0.0, dtype=F64))
        lp = _logit(p_i)
        logit_pi = torch.where(
            torch.isfinite(lp),
            lp,
            torch.where(
                _isclose(psi[i], torch.tensor(0.0, dtype=F64)) & _isclose(rho[i], torch.tensor(0.0, dtype=F64)),
                torch.where(
                    lam[i] < mu[i],
                    logit_pi1 - torch.log(pi1) - q - (lam[i] - torch.exp(q) * mu[i]) * (1.0 - pi1) / (lam[i] - mu[i]),
                    torch.log(mu[i] / (lam[i] - mu[i]).abs()),
                ),
                lp,
            ),
        )
        pi1, logit_pi1 = p_i, logit_pi
        B[i] = B_i
    p1_0, logit_p1_0 = pi1, logit_pi1
    B = torch.stack(B)

    x = xs[n:]
    y = xs[:n]
    Ix = torch.searchsorted(times.detach(), x.detach(), right=True)
    Iy = torch.searchsorted(times.detach(), y.detach(), right=True)

    def log_q(i, t):
        # bdsky.py:136-143 (lines 144-149 are unreachable)
        A_i, B_i = _take(A, i - 1), _take(B, i - 1)
        t_i = _take(times, i)
        Adt = A_i * (t - t_i)
        return math.log(4) + Adt - 2 * torch.log((torch.exp(Adt) * (1 - B_i) + (1 + B_i)).abs())

    # bdsky.py:152-158
    node_is_sample = n <= np.arange(2 * n - 1)
    ev = order_events(times.detach().numpy(), xs.detach().numpy(), node_is_sample)
    idx = np.searchsorted(ev[:, 0], times.detach().numpy()[1:-1], side="left")
    n_i = torch.as_tensor(np.append(1 + ev[idx, 1], 0.0), dtype=F64)

    # bdsky.py:161-163
    near_t_i = _isclose(xs[:n, None], times[None, 1:]) & (rho[None] > 0)
    N_i = near_t_i.sum(0)
    sampled_leaves = ~(near_t_i.max(1).values)

    one = torch.tensor(1, dtype=torch.long)
    l1 = log_q(one, times[0])
    if condition_on_survival:
        l1_0 = torch.log1p(-p1_0)
        l1_1 = torch.where(logit_p1_0 > 10, -logit_p1_0, -torch.log1p(torch.exp(logit_p1_0)))
        l1_2 = torch.where(torch.isfinite(l1_0), l1_0, l1_1)
        l1 = l1 - l1_2
    loglik = l1

    l2 = log_q(Ix, x) + _take(torch.log(lam), Ix - 1)
    loglik = loglik + l2.sum()

    sl = sampled_leaves.to(F64)
    l3 = _xlogy(sampled_leaves, _take(psi, Iy - 1)) - sl * log_q(Iy, y)
    loglik = loglik + l3.sum()

    ii = torch.arange(1, m)
    log_qi1_ti = log_q(ii + 1, times[1:-1])
    l41 = _xlog1py(n_i, -rho)
    l42 = n_i[:-1] * log_qi1_ti
    loglik = loglik + l41.sum() + l42.sum()

    l5 = _xlogy(N_i, rho)
    loglik = loglik + l5.sum()
    if return_parts:
        return loglik, dict(l1=l1, l2=l2, l3=l3, l41=l41, l42=l42, l5=l5, A=A, B=B, p1_0=p1_0,
                            logit_p1_0=logit_p1_0, n_i=n_i, N_i=N_i, Ix=Ix, Iy=Iy)
    return loglik


def bdsky_transform(params):
    """bdsky.py:219-227."""
    lam = params["R"] * params["delta"]
    psi = params["s"] * params["delta"]
    mu = params["delta"] - psi
    return lam, psi, mu, params["rho"]


def height_transform(td, root_height, proportions):
    """tree_data.py:287-367 on torch tensors (differentiable in root_height, proportions)."""
    n, N = td.n, td.N
    lst = torch.as_tensor(td.lower_sampling_times, dtype=F64)
    h = [None] * N
    for i in range(n):
        h[i] = lst[i]
    h[N - 1] = root_height + float(td.sample_times.max())
    if N - n > 1:
        for i in td.preorder[1:]:
            if i >= n:
                p = proportions[i - n]
                h[i] = (1.0 - p) * lst[i] + p * h[td.child_parent[i]]
    return torch.stack([hh.reshape(()) for hh in h])


class _PruneData(torch.autograd.Function):
    """bdsky.py:331-343 with prune.py:134-150's custom derivative: value = sum_s counts[s]*ll_s,
    d/d(branch_lengths) by eq. (9); no derivative w.r.t. Q, pi, tips (prune.py:149)."""

    @staticmethod
    def forward(ctx, bl, Q, partials, counts, td):
        tot, g, _ = oprune.data_loglik_and_grad(bl.detach().numpy(), Q, partials, counts, td)
        ctx.g = torch.as_tensor(g, dtype=F64)
        return torch.tensor(tot, dtype=F64)

    @staticmethod
    def backward(ctx, gout):
        return gout * ctx.g, None, None, None, None


def loglik(params, td, tip_data, Q, c=(True, True, True), condition_on_survival=True,
           equidistant_intervals=True, params_prior_loglik=lambda p: 0.0, return_parts=False):
    """bdsky.py:268-345.  ``params`` maps names to torch fp64 tensors (1-D, as in the reference)."""
    assert len(params["proportions"]) == td.n - 2
    assert len(params["R"]) == len(params["s"]) == len(params["delta"])
    assert params["root_proportion"].ndim == 1 and params["root_proportion"].numel() == 1
    ll = torch.tensor(0.0, dtype=F64)
    parts = {}
    if c[0]:
        ll = ll + params_prior_loglik(params)
    lam, psi, mu, rho = bdsky_transform(params)
    root_height = params["root_proportion"][0] * (
        params["origin"][0] + params["origin_start"][0] - float(td.sample_times.max())
    )
    node_heights = height_transform(td, root_height, params["proportions"])
    cp = torch.as_tensor(td.child_parent[:-1], dtype=torch.long)
    branch_lengths = node_heights[cp] - node_heights[:-1]
    tm = params["origin"][0] + params["origin_start"][0]
    if equidistant_intervals:
        m = len(params["R"])
        # jnp.linspace(0, tm, m+1): start*(1-i/m) + stop*(i/m) for i<m, endpoint = stop exactly.
        times = tm * (torch.arange(m + 1, dtype=F64) / m)
    else:
        times = tm - params["grid"]
        times = torch.cat([torch.zeros(1, dtype=F64), times[1:]])
    xs = tm - node_heights
    if c[1]:
        tl = tree_loglik(lam, psi, mu, rho, xs, times, td.n, condition_on_survival)
        parts["tree"] = tl
        ll = ll + tl
    if c[2]:
        dl = _PruneData.apply(branch_lengths * params["clock_rate"][0], Q, tip_data.partials, tip_data.counts, td)
        parts["data"] = dl
        ll = ll + dl
    if return_parts:
        parts.update(node_heights=node_heights, branch_lengths=branch_lengths, xs=xs, times=times)
        return ll, parts
    return ll


def unpack(samples):
    """optim.py:19-32 on a dict of [K, ...] torch tensors."""
    out = {}
    for k, v in samples.items():
        if isinstance(v, dict):
            out |= v
        else:
            out[k] = v
    out["rho"] = torch.cat([torch.zeros_like(out["delta"][:, :-1]), out["rho_m"]], dim=1)
    return out
