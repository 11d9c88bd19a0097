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
    def forward(ctx, mu, log_sigma, eps, unit_tree, rows):
        U, d = eps.shape
        ctx.rows = rows
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
        if ctx.rows:  # one parameter row per unit: the per-unit gradients are the result
            return dmu_u, dls_u, None, None, None
        # fixed-order segmented sum over each tree's units (no atomics: results are bitwise reproducible)
        order = torch.argsort(unit_tree.long(), stable=True)
        sorted_tree = unit_tree.long()[order]
        trees, counts = torch.unique_consecutive(sorted_tree, return_counts=True)
        dmu, dls = torch.zeros_like(mu), torch.zeros_like(log_sigma)
        if counts.numel() and bool((counts == counts[0]).all()):
            k = int(counts[0])
            dmu[trees] = dmu_u[order].view(-1, k, d).sum(1)
            dls[trees] = dls_u[order].view(-1, k, d).sum(1)
        else:  # ragged: one pass per rank-within-tree, still in a fixed order
            start = torch.cumsum(counts, 0) - counts
            rank = torch.arange(U, device=eps.device) - start.repeat_interleave(counts)
            for r in range(int(counts.max())):
                sel = order[rank == r]
                dmu[unit_tree.long()[sel]] += dmu_u[sel]
                dls[unit_tree.long()[sel]] += dls_u[sel]
        return dmu, dls, None, None, None


def local_flow(mu: torch.Tensor, log_sigma: torch.Tensor, eps: torch.Tensor, unit_tree: torch.Tensor):
    """mu, log_sigma [T, d] (per-tree DiagonalAffine parameters), eps [U, d] standard-normal base draws,
    unit_tree [U] int32 -> (x [U, d] in (0, 1), log q(x) [U]); differentiable in mu and log_sigma."""
    for t in (mu, log_sigma, eps):
        if t.dtype != torch.float64 or not t.is_cuda:
            raise ValueError("local_flow expects CUDA float64 tensors")
    assert mu.shape == log_sigma.shape and eps.shape[1] == mu.shape[1] and unit_tree.shape[0] == eps.shape[0]
    return _LocalFlow.apply(mu, log_sigma, eps, unit_tree, False)


def local_flow_rows(mu_u: torch.Tensor, log_sigma_u: torch.Tensor, eps: torch.Tensor):
    """Same map with one parameter row per unit (mu_u, log_sigma_u [U, d], e.g. an ``expand`` of the per-tree rows, whose
    backward is a fixed-order sum): the gradients come back per unit and the caller's autograd graph reduces them."""
    for t in (mu_u, log_sigma_u, eps):
        if t.dtype != torch.float64 or not t.is_cuda:
            raise ValueError("local_flow_rows expects CUDA float64 tensors")
    assert mu_u.shape == log_sigma_u.shape == eps.shape
    rows = torch.arange(eps.shape[0], dtype=torch.int32, device=eps.device)
    return _LocalFlow.apply(mu_u, log_sigma_u, eps, rows, True)
