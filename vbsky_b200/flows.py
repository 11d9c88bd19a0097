"""The reference's per-tree (local) variational family evaluated on the device: reparameterised
samples of ``proportions`` / ``root_proportion`` and their log-density for the ELBO's entropy term
(fasta.py:378-384: ``Transform(td.n - 2, z1)``, ``Transform(1, z1)`` with
``z1 = Compose(DiagonalAffine, ZeroOne)``; optim.py:54-65).  Batched over (tree, sample) units."""
import ctypes as C

import torch

from . import _lib


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


class _LocalFlow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, log_sigma, eps, unit_tree):
        U, d = eps.shape
        mu, log_sigma, eps = mu.contiguous(), log_sigma.contiguous(), eps.contiguous()
        x = torch.empty_like(eps)
        logq = torch.empty(U, dtype=torch.float64, device=eps.device)
        with torch.cuda.device(eps.device):
            st = C.c_void_p(torch.cuda.current_stream(eps.device).cuda_stream)
            _lib.check(_lib.lib.vbsky_local_flow_sample(U, d, _p(unit_tree), _p(mu), _p(log_sigma), _p(eps), _p(x), _p(logq), st))
        ctx.save_for_backward(mu, log_sigma, eps, unit_tree)
        return x, logq

    @staticmethod
    def backward(ctx, gx, gq):
        mu, log_sigma, eps, unit_tree = ctx.saved_tensors
        U, d = eps.shape
        gx = None if gx is None else gx.contiguous()
        gq = None if gq is None else gq.contiguous()
        dmu_u = torch.empty_like(eps)
        dls_u = torch.empty_like(eps)
        with torch.cuda.device(eps.device):
            st = C.c_void_p(torch.cuda.current_stream(eps.device).cuda_stream)
            _lib.check(_lib.lib.vbsky_local_flow_backward(U, d, _p(unit_tree), _p(mu), _p(log_sigma), _p(eps), _p(gx), _p(gq),
                                                          _p(dmu_u), _p(dls_u), st))
        idx = unit_tree.long()
        dmu = torch.zeros_like(mu).index_add_(0, idx, dmu_u)
        dls = torch.zeros_like(log_sigma).index_add_(0, idx, dls_u)
        return dmu, dls, None, None


def local_flow(mu: torch.Tensor, log_sigma: torch.Tensor, eps: torch.Tensor, unit_tree: torch.Tensor):
    """mu, log_sigma [T, d] (per-tree DiagonalAffine parameters), eps [U, d] standard-normal base draws,
    unit_tree [U] int32 -> (x [U, d] in (0, 1), log q(x) [U]); differentiable in mu and log_sigma."""
    for t in (mu, log_sigma, eps):
        if t.dtype != torch.float64 or not t.is_cuda:
            raise ValueError("local_flow expects CUDA float64 tensors")
    assert mu.shape == log_sigma.shape and eps.shape[1] == mu.shape[1] and unit_tree.shape[0] == eps.shape[0]
    return _LocalFlow.apply(mu, log_sigma, eps, unit_tree)
