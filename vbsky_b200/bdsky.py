"""Model log-density of the reference's ``bdsky.py`` evaluated by the CUDA path.

``tree_loglik`` mirrors ``_tree_loglik(lam, psi, mu, rho, xs, times, td, ...)`` (bdsky.py:18-216) and
``loglik`` mirrors ``bdsky.loglik(params, tr_d, tp_d, Q, c, ...)`` (bdsky.py:268-345), both batched
over a leading unit axis (the reference vmaps over Monte-Carlo samples, optim.py:72) and wrapped in
``torch.autograd.Function`` so that the surrounding Python (flows, priors, optimiser) differentiates
through them exactly as it does through the JAX originals.  All arithmetic happens in the kernels
behind ``vbsky_bdsky_value_and_grad`` / ``vbsky_loglik_value_and_grad``; there is no CPU path.
"""
import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib

TERM_TREE, TERM_DATA, CONDITION_ON_SURVIVAL = 1, 2, 4
PARAM_KEYS = ("proportions", "root_proportion", "origin", "origin_start", "clock_rate", "R", "delta", "s", "rho", "grid")
GRAD_KEYS = ("proportions", "root_proportion", "origin", "clock_rate", "R", "delta", "s", "rho", "grid")


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _TreeLoglik(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lam, psi, mu, rho, xs, times, n, condition_on_survival):
        U, m = lam.shape
        dev = lam.device
        args = [t.contiguous() for t in (lam, psi, mu, rho, xs, times)]
        lp = torch.empty(U, dtype=torch.float64, device=dev)
        g = [torch.empty_like(t) for t in args]
        with torch.cuda.device(dev):
            nbytes = _lib.lib.vbsky_bdsky_workspace_bytes(U, n)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(_lib.lib.vbsky_bdsky_value_and_grad(
                U, n, m, *[_p(t) for t in args], int(condition_on_survival), _p(lp),
                _p(g[0]), _p(g[1]), _p(g[2]), _p(g[3]), _p(g[4]), _p(g[5]), _p(ws), nbytes, _stream(dev)))
        ctx.save_for_backward(*g)
        return lp

    @staticmethod
    def backward(ctx, gout):
        g = ctx.saved_tensors
        return tuple(gout[:, None] * t for t in g) + (None, None)


def tree_loglik(lam, psi, mu, rho, xs, times, n: int, condition_on_survival: bool = True) -> torch.Tensor:
    """BDSKY log-density (Stadler et al. 2013, Thm 1) for U units: lam, psi, mu, rho [U, m], xs
    [U, 2n-1] forward node times (tips first), times [U, m+1].  CUDA fp64 tensors; returns [U]."""
    for t in (lam, psi, mu, rho, xs, times):
        if t.dtype != torch.float64 or not t.is_cuda:
            raise ValueError("tree_loglik expects CUDA float64 tensors")
    m = lam.shape[1]
    assert lam.shape == psi.shape == mu.shape == rho.shape  # bdsky.py:44
    assert times.shape[1] == m + 1  # bdsky.py:45
    assert xs.shape[1] == 2 * n - 1
    return _TreeLoglik.apply(lam, psi, mu, rho, xs, times, n, condition_on_survival)


class _Loglik(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ens, unit_tree, m, flags, clean, *vals):
        params = dict(zip(PARAM_KEYS, vals))
        U = params["root_proportion"].shape[0]
        dev = ens.device
        need_grad = any(v is not None and v.requires_grad for v in vals)
        cin = _lib.Params()
        keep = []
        for k in PARAM_KEYS:
            v = params[k]
            if v is not None:
                v = v.contiguous()
                keep.append(v)
            setattr(cin, k, 0 if v is None else v.data_ptr())
        ll = torch.empty(U, dtype=torch.float64, device=dev)
        lt = torch.empty(U, dtype=torch.float64, device=dev)
        ld = torch.empty(U, dtype=torch.float64, device=dev)
        grads = {}
        cg = None
        if need_grad:
            cg = _lib.ParamGrads()
            for k in GRAD_KEYS:
                v = params[k]
                grads[k] = None if v is None else torch.empty_like(v, memory_format=torch.contiguous_format)
                setattr(cg, k, 0 if grads[k] is None else grads[k].data_ptr())
        with torch.cuda.device(dev):
            nbytes = _lib.lib.vbsky_loglik_workspace_bytes(ens._layout, ens._tips, U, m)
            if nbytes == 0:
                raise _lib.VbskyError(-1, _lib.last_error())
            ws = ens._workspace(nbytes)
            _lib.check(_lib.lib.vbsky_loglik_value_and_grad(
                ens._layout, ens._tips, U, _p(unit_tree), m, C.byref(cin), C.byref(ens._model), flags,
                _p(ll), _p(lt), _p(ld), None if cg is None else C.byref(cg), _p(ws), ws.numel(), _stream(dev)))
        if clean:  # fasta.py:421: nan_to_num on every single-sample (value, gradient) before they are averaged
            ll = torch.nan_to_num(ll)
            grads = {k: (None if v is None else torch.nan_to_num(v)) for k, v in grads.items()}
        ctx.grads = grads
        ctx.mark_non_differentiable(lt, ld)
        return ll, lt, ld

    @staticmethod
    def backward(ctx, gout, _gt, _gd):
        g = ctx.grads
        out = []
        for k in PARAM_KEYS:
            src = "origin" if k == "origin_start" else k
            v = g.get(src)
            if v is None:
                out.append(None)
            else:
                out.append(gout.reshape((-1,) + (1,) * (v.dim() - 1)) * v)
        return (None, None, None, None, None) + tuple(out)


def loglik(params: Dict[str, torch.Tensor], ens, unit_tree: torch.Tensor, c=(True, True, True),
           condition_on_survival: bool = True, params_prior_loglik=None, return_parts: bool = False, clean: bool = False,
           equidistant_intervals: bool = True):
    """bdsky.loglik for U units.  ``params`` holds CUDA fp64 tensors with a leading unit axis:
    proportions [U, n-2], root_proportion/origin/origin_start/clock_rate [U] (or [U, 1]), R/delta/s/rho
    [U, m]; ``ens`` is the device-resident ``Ensemble``; ``unit_tree`` [U] int32 picks each unit's tree.
    ``c = (prior, tree, data)`` switches the three terms as in the reference (bdsky.py:288,323,331);
    the prior term is the user's ``params_prior_loglik(params)`` -> [U] evaluated in torch.
    ``clean=True`` applies ``nan_to_num`` to every unit's value and gradient rows, which is what the reference's ``step``
    does to each single-sample (f, g) before averaging (fasta.py:421), so one non-finite draw is dropped, not spread.
    ``equidistant_intervals=False`` takes the prespecified grid ``params["grid"]`` [U, m+1] (bdsky.py:317-319:
    ``times = tm - grid`` with ``times[0] = 0``) instead of ``linspace(0, tm, m+1)``; it is differentiable too."""
    U = unit_tree.shape[0]
    n = ens.n
    flat = {}
    for k in PARAM_KEYS:
        v = params.get(k) if (k != "grid" or not equidistant_intervals) else None
        if k == "grid" and not equidistant_intervals and v is None:
            raise ValueError("equidistant_intervals=False needs params['grid'] [U, m+1]")
        if v is None:
            flat[k] = None
            continue
        if k in ("root_proportion", "origin", "origin_start", "clock_rate"):
            v = v.reshape(U)
        flat[k] = v
    assert flat["proportions"].shape == (U, n - 2)  # bdsky.py:282
    m = 0 if flat["R"] is None else flat["R"].shape[1]
    if c[1]:
        assert flat["R"].shape == flat["s"].shape == flat["delta"].shape == flat["rho"].shape  # bdsky.py:283
    flags = (TERM_TREE if c[1] else 0) | (TERM_DATA if c[2] else 0) | (CONDITION_ON_SURVIVAL if condition_on_survival else 0)
    ll, lt, ld = _Loglik.apply(ens, unit_tree, m, flags, bool(clean), *[flat[k] for k in PARAM_KEYS])
    if c[0] and params_prior_loglik is not None:
        ll = ll + params_prior_loglik(params)
    if return_parts:
        return ll, {"tree": lt, "data": ld}
    return ll


def bdsky_transform(params):
    """R, delta, s -> (lam, psi, mu, rho) as bdsky.py:219-227 (the kernels do this internally)."""
    lam = params["R"] * params["delta"]
    psi = params["s"] * params["delta"]
    return lam, psi, params["delta"] - psi, params["rho"]


def lognorm_logpdf(log_x, mu, sigma):
    """Log-normal log-density in terms of log x (bdsky.py:230-231); used by callers' parameter priors."""
    import math

    sigma = torch.as_tensor(sigma, dtype=log_x.dtype, device=log_x.device)
    z = (log_x - mu) / sigma
    return -0.5 * z * z - 0.5 * math.log(2.0 * math.pi) - log_x - torch.log(sigma)
