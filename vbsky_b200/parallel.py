"""Multi-GPU plumbing of the ensemble ELBO: the (tree, sample) units are independent up to a final
sum, so they are dealt round-robin to the ranks (one process per GPU) and the only exchange per
evaluation is one small all-reduce of [ELBO value, gradient w.r.t. the global parameters]; gradients of
per-tree (local) parameters stay on the rank that owns the unit.  torch.distributed provides the
collective (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

GLOBAL_KEYS = ("origin", "clock_rate", "R", "delta", "s", "rho")


def shard_units(T: int, K: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """Units (tree t, sample k) owned by ``rank``: unit index u = t*K + k, taken round-robin (u % world == rank)
    so that T*K = 1000 units split 125/125/... over 8 ranks instead of 13/12 trees."""
    return [(u // K, u % K) for u in range(rank, T * K, world)]


def pack_global(ll: torch.Tensor, grads: Dict[str, torch.Tensor], K: int, out: torch.Tensor = None) -> torch.Tensor:
    """[sum_u ll_u / K, sum_u d/d(global param) / K ...] as one flat fp64 vector (the all-reduce payload)."""
    parts = [ll.sum().reshape(1)]
    for k in GLOBAL_KEYS:
        g = grads[k]
        parts.append(g.sum(0).reshape(-1) if g.dim() > 1 else g.sum().reshape(1))
    v = torch.cat(parts) / K
    if out is not None:
        out.copy_(v)
        return out
    return v


def allreduce_elbo(vec: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the packed vector over ranks in place (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def unpack_global(vec: torch.Tensor, m: int) -> Dict[str, torch.Tensor]:
    out = {"elbo": vec[0], "origin": vec[1], "clock_rate": vec[2]}
    for i, k in enumerate(("R", "delta", "s", "rho")):
        out[k] = vec[3 + i * m: 3 + (i + 1) * m]
    return out
