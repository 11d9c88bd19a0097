"""JAX side of the drop-in boundary: ``jax.ffi`` custom calls into libvbsky_b200 with one ``custom_vjp`` per call, so
that the reference's optimiser loop and reparameterised sampling (optim.py, fasta.py) run unchanged on top of them.

    prune.prune_loglik vmapped over sites x counts (prune.py:99, bdsky.py:331-343)   ->  data_loglik
    bdsky._tree_loglik (bdsky.py:18)                                                  ->  tree_loglik
    TreeData.height_transform + the glue of bdsky.loglik (tree_data.py:287, bdsky.py:292-322) -> heights
    bdsky.loglik under optim.loss's vmap over samples (bdsky.py:268, optim.py:72)     ->  loglik

``custom_vjp`` (not the reference's ``custom_jvp``, prune.py:99): an opaque FFI call cannot be transposed by JAX, so every
forward pass also returns the analytic gradient as its residual and the backward pass is a product with the cotangent.

Requires jax >= 0.4.31 (``jax.ffi``) and ``libvbsky_ffi.so`` built with ``make -C vbsky_b200/csrc ffi XLA_FFI_INCLUDE=...``
(csrc/ffi/vbsky_ffi.cc).  JAX is not installable in this repository's image, so this module is exercised only where
JAX exists; tests/test_ffi_source.py checks the handler source against the C header, and the same C entry points are
tested through the ctypes/PyTorch harness (vbsky_b200/bdsky.py).  Nothing in the package imports this module.
"""
import ctypes as C
import os
from functools import partial
from typing import NamedTuple

import numpy as np

try:
    import jax
    import jax.numpy as jnp
except ImportError as e:  # pragma: no cover - the repository image has no JAX
    raise ImportError("vbsky_b200.jax_ffi needs JAX (>= 0.4.31); use vbsky_b200.bdsky (ctypes/PyTorch harness) otherwise") from e

from . import _lib

_HERE = os.path.dirname(os.path.abspath(__file__))
_FFI_PATH = os.environ.get("VBSKY_FFI_LIB") or os.path.join(_HERE, "libvbsky_ffi.so")
if not os.path.exists(_FFI_PATH):
    raise ImportError(f"{_FFI_PATH} not found: build it with `make -C vbsky_b200/csrc ffi XLA_FFI_INCLUDE=$(python -c 'import jax.ffi; "
                      "print(jax.ffi.include_dir())')`")
_ffi = C.CDLL(_FFI_PATH)

TARGETS = {
    "vbsky_prune": "VbskyPrune", "vbsky_bdsky": "VbskyBdsky", "vbsky_heights_forward": "VbskyHeightsForward",
    "vbsky_heights_backward": "VbskyHeightsBackward", "vbsky_loglik": "VbskyLoglik", "vbsky_reduce_elbo": "VbskyReduceElbo",
}
for _name, _sym in TARGETS.items():
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_ffi, _sym)), platform="CUDA")

PARAM_KEYS = ("proportions", "root_proportion", "origin", "origin_start", "clock_rate", "R", "delta", "s", "rho")
GRAD_KEYS = ("proportions", "root_proportion", "origin", "clock_rate", "R", "delta", "s", "rho")
F64 = jnp.float64


class Handles(NamedTuple):
    """Static (hashable) description of a device-resident ensemble: raw handle addresses of ``vbsky_b200.ensemble.Ensemble``
    plus the substitution model as tuples (SubstitutionModel of substitution.py:51-55)."""
    layout: int
    tips: int
    n: int
    S: int
    pi: tuple
    P: tuple
    d: tuple
    Pinv: tuple

    @classmethod
    def of(cls, ens):
        Q = ens.Q
        flat = lambda a: tuple(np.asarray(a, dtype=np.float64).reshape(-1).tolist())
        return cls(int(ens._layout.value), int(ens._tips.value), int(ens.n), int(ens.S), flat(Q.pi), flat(Q.P), flat(Q.d), flat(Q.Pinv))

    def model_attrs(self):
        return dict(pi=np.array(self.pi), P=np.array(self.P), d=np.array(self.d), Pinv=np.array(self.Pinv))


def _sds(shape, dtype=F64):
    return jax.ShapeDtypeStruct(tuple(int(x) for x in shape), dtype)


def _ws(nbytes):
    if nbytes == 0:
        raise _lib.VbskyError(-1, _lib.last_error())
    return _sds((int(nbytes),), jnp.uint8)


# ------------------------------------------------------------------------------------------------------------ data term
@partial(jax.custom_vjp, nondiff_argnums=(2,))
def data_loglik(bl_scaled, unit_tree, h: Handles):
    """sum_s counts[s] * prune_loglik(bl_scaled[u], Q, partials[s], td) for every unit u (bdsky.py:331-343).
    bl_scaled [U, 2n-2] = branch_lengths * clock_rate; unit_tree [U] int32."""
    return _data_fwd(bl_scaled, unit_tree, h)[0]


def _data_fwd(bl_scaled, unit_tree, h):
    U, nb = bl_scaled.shape
    nbytes = _lib.lib.vbsky_prune_workspace_bytes(C.c_void_p(h.layout), C.c_void_p(h.tips), U, 1)
    ll, dll, _ = jax.ffi.ffi_call("vbsky_prune", (_sds((U,)), _sds((U, nb)), _ws(nbytes)))(
        unit_tree, bl_scaled, layout=np.int64(h.layout), tips=np.int64(h.tips), **h.model_attrs())
    return ll, (dll, unit_tree)


def _data_bwd(h, res, g):
    dll, unit_tree = res
    return g[:, None] * dll, None  # eq. (9) of the paper (prune.py:146-148); Q / pi / tip tangents unsupported as in prune.py:149


data_loglik.defvjp(_data_fwd, _data_bwd)


# ----------------------------------------------------------------------------------------------------------- tree prior
@partial(jax.custom_vjp, nondiff_argnums=(6, 7))
def tree_loglik(lam, psi, mu, rho, xs, times, n: int, condition_on_survival: bool = True):
    """bdsky._tree_loglik (bdsky.py:18-216) for U units: lam, psi, mu, rho [U, m], xs [U, 2n-1], times [U, m+1] -> [U]."""
    return _tree_fwd(lam, psi, mu, rho, xs, times, n, condition_on_survival)[0]


def _tree_fwd(lam, psi, mu, rho, xs, times, n, condition_on_survival):
    U, m = lam.shape
    nbytes = _lib.lib.vbsky_bdsky_workspace_bytes(U, n)
    outs = jax.ffi.ffi_call(
        "vbsky_bdsky",
        (_sds((U,)), _sds((U, m)), _sds((U, m)), _sds((U, m)), _sds((U, m)), _sds((U, 2 * n - 1)), _sds((U, m + 1)), _ws(nbytes)),
    )(lam, psi, mu, rho, xs, times, n=np.int32(n), condition_on_survival=np.int32(bool(condition_on_survival)))
    return outs[0], outs[1:7]


def _tree_bwd(n, condition_on_survival, res, g):
    return tuple(g[:, None] * r for r in res)


tree_loglik.defvjp(_tree_fwd, _tree_bwd)


# --------------------------------------------------------------------------------------------------------- height chain
@partial(jax.custom_vjp, nondiff_argnums=(2,))
def heights(params, unit_tree, h: Handles):
    """TreeData.height_transform and the glue of bdsky.loglik (tree_data.py:287-367, bdsky.py:292-322) for U units.
    ``params``: dict of PARAM_KEYS -> [U, ...].  Returns (heights [U, 2n-1], bl_scaled [U, 2n-2], xs [U, 2n-1],
    times [U, m+1], lam, psi, mu [U, m])."""
    return _heights_fwd(params, unit_tree, h)[0]


def _heights_call(params, unit_tree, h):
    U, m = params["R"].shape
    N = 2 * h.n - 1
    return jax.ffi.ffi_call(
        "vbsky_heights_forward",
        (_sds((U, N)), _sds((U, N - 1)), _sds((U, N)), _sds((U, m + 1)), _sds((U, m)), _sds((U, m)), _sds((U, m))),
    )(unit_tree, *[params[k] for k in PARAM_KEYS], layout=np.int64(h.layout))


def _heights_fwd(params, unit_tree, h):
    out = _heights_call(params, unit_tree, h)
    return tuple(out), (params, unit_tree, out[0])


def _heights_bwd(h, res, cot):
    params, unit_tree, hts = res
    _, d_bl, d_xs, d_times, d_lam, d_psi, d_mu = cot  # a cotangent on `heights` itself enters through xs / bl_scaled in bdsky.loglik
    U, m = params["R"].shape
    n = h.n
    g = jax.ffi.ffi_call(
        "vbsky_heights_backward",
        (_sds((U, n - 2)), _sds((U,)), _sds((U,)), _sds((U,)), _sds((U, m)), _sds((U, m)), _sds((U, m)), _ws(U * (2 * n - 1) * 8)),
    )(unit_tree, *[params[k] for k in PARAM_KEYS], hts, d_bl, d_xs, d_times, d_lam, d_psi, d_mu, layout=np.int64(h.layout))
    gp = dict(zip(("proportions", "root_proportion", "origin", "clock_rate", "R", "delta", "s"), g[:7]))
    gp["origin_start"] = gp["origin"]  # both enter as a sum (bdsky.py:299, 312)
    gp["rho"] = jnp.zeros_like(params["rho"])
    return {k: gp[k].reshape(params[k].shape) for k in PARAM_KEYS}, None


heights.defvjp(_heights_fwd, _heights_bwd)


# --------------------------------------------------------------------------------------------------------- fused loglik
TERM_TREE, TERM_DATA, CONDITION_ON_SURVIVAL = 1, 2, 4


@partial(jax.custom_vjp, nondiff_argnums=(2, 3))
def loglik(params, unit_tree, h: Handles, flags: int = TERM_TREE | TERM_DATA | CONDITION_ON_SURVIVAL):
    """bdsky.loglik (bdsky.py:268-345) without the user prior term c[0], for U units in one call: the unit axis is the
    sample axis that optim.loss vmaps over (optim.py:72) times the trees of the ensemble."""
    return _loglik_fwd(params, unit_tree, h, flags)[0]


def _loglik_fwd(params, unit_tree, h, flags):
    U, m = params["R"].shape
    nbytes = _lib.lib.vbsky_loglik_workspace_bytes(C.c_void_p(h.layout), C.c_void_p(h.tips), U, m)
    shapes = (_sds((U,)), _sds((U,)), _sds((U,))) + tuple(_sds(params[k].shape) for k in GRAD_KEYS) + (_ws(nbytes),)
    outs = jax.ffi.ffi_call("vbsky_loglik", shapes)(
        unit_tree, *[params[k] for k in PARAM_KEYS], layout=np.int64(h.layout), tips=np.int64(h.tips), flags=np.int32(flags),
        **h.model_attrs())
    return outs[0], dict(zip(GRAD_KEYS, outs[3:3 + len(GRAD_KEYS)]))


def _loglik_bwd(h, flags, grads, g):
    out = {}
    for k in PARAM_KEYS:
        v = grads["origin" if k == "origin_start" else k]
        out[k] = g.reshape((-1,) + (1,) * (v.ndim - 1)) * v
    return out, None


loglik.defvjp(_loglik_fwd, _loglik_bwd)


def reduce_elbo(unit_tree, ll, grads, T: int, n: int, K: int):
    """vbsky_reduce_elbo: the all-reduce payload [ELBO, d/d global, d/d per-tree local] / K of one evaluation;
    follow with ``jax.lax.psum`` over the device axis."""
    m = grads["R"].shape[1]
    return jax.ffi.ffi_call("vbsky_reduce_elbo", _sds((3 + 4 * m + T * (n - 1),)))(
        unit_tree, ll, *[grads[k] for k in GRAD_KEYS], T=np.int32(T), n=np.int32(n), scale=np.float64(1.0 / K))
