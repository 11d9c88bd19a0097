"""``prune_loglik`` with the reference's signature (/root/reference/vbsky/prune.py:99-131), evaluated
by the CUDA pruning kernel.  The reference vmaps this function over site patterns
(bdsky.py:333); here ``tip_partials`` may already carry that leading axis ([S, n, 4]).
The derivative w.r.t. ``branch_lengths`` is the kernel's analytic gradient (prune.py:134-150);
Q, pi and the tip partials get no gradient, as in the reference (prune.py:149)."""
import numpy as np
import torch

from .ensemble import Ensemble
from .substitution import SubstitutionModel, partials_to_codes
from .tips import PackedTips
from .tree_data import TreeData


class _PruneSites(torch.autograd.Function):
    """Per-pattern log-likelihoods of one tree; the cotangent g [S] is pulled back with one more kernel call in
    which the pattern weights are g (d/d bl of sum_s g[s] * ll_s is exactly what the kernel accumulates)."""

    @staticmethod
    def forward(ctx, bl, ens):
        unit_tree = torch.zeros(1, dtype=torch.int32, device=bl.device)
        _, _, site = ens.prune(bl[None].contiguous(), unit_tree, want_grad=False, want_site_ll=True)
        ctx.ens = ens
        ctx.save_for_backward(bl)
        return site[0]

    @staticmethod
    def backward(ctx, gout):
        (bl,) = ctx.saved_tensors
        ens = ctx.ens
        ens.set_counts(gout.detach().double().cpu().numpy()[None])
        unit_tree = torch.zeros(1, dtype=torch.int32, device=bl.device)
        _, dll, _ = ens.prune(bl[None].contiguous(), unit_tree, want_grad=True)
        return dll[0], None


def prune_loglik(branch_lengths, Q: SubstitutionModel, tip_partials, td: TreeData, rescale: bool = True, dbg: bool = False):
    """Log-likelihood of each site pattern under the tree/substitution model.

    branch_lengths [2n-2] (torch CUDA fp64 tensor, or array-like), tip_partials [n,4] or [S,n,4]
    0/1 indicator partials.  Returns a CUDA fp64 tensor of shape [] or [S].  ``rescale`` must stay
    True (the kernel always rescales; prune.py's rescale=False path is never used by the reference)."""
    if not rescale:
        raise NotImplementedError("rescale=False is not supported by the CUDA path")
    tp = np.asarray(tip_partials)
    single = tp.ndim == 2
    if single:
        tp = tp[None]
    assert tp.ndim == 3 and tp.shape[1:] == (td.n, 4), "tip_partials must be [n,4] or [S,n,4]"
    if not isinstance(branch_lengths, torch.Tensor):
        branch_lengths = torch.as_tensor(np.asarray(branch_lengths, dtype=np.float64))
    bl = branch_lengths.to(device="cuda", dtype=torch.float64)
    assert bl.ndim == 1 and bl.shape[0] == 2 * (td.n - 1)
    patterns = partials_to_codes(tp)
    ens = Ensemble([td], [PackedTips(patterns, np.ones(len(patterns)))], Q, device=bl.device)
    out = _PruneSites.apply(bl, ens)
    return out[0] if single else out
