"""Multi-GPU plumbing of the ensemble ELBO: the (tree, sample) units are independent up to a final sum, so they are
dealt round-robin to the ranks (one process per GPU) and the only exchange per evaluation is ONE all-reduce of
[ELBO value, gradient w.r.t. the global parameters, gradient w.r.t. every tree's local parameters].

The K samples of one tree land on different ranks (that is what balances 1000 units over 8 GPUs), so every rank holds
only a partial sum of a tree's local gradient: the local block is therefore part of the all-reduce payload
(3 + 4m + T(n-1) doubles; 160 KB at 100 trees x 200 tips -- latency-bound over NVLink), after which every rank holds
the complete gradient and applies the same optimiser update.  The payload is built by one kernel
(``vbsky_reduce_elbo``, fixed summation order).  torch.distributed provides the collective (NCCL over NVLink on GPUs;
gloo in the CPU tests of the host logic)."""
import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

GLOBAL_KEYS = ("origin", "clock_rate", "R", "delta", "s", "rho")


def shard_units(T: int, K: int, rank: int, world: int) -> List[Tuple[int, int]]:
    """Units (tree t, sample k) owned by ``rank``: unit index u = t*K + k, taken round-robin (u % world == rank)
    so that T*K = 1000 units split 125/125/... over 8 ranks instead of 13/12 trees.  The returned list is ordered by
    unit index, i.e. grouped by tree (what ``reduce_elbo`` requires)."""
    return [(u // K, u % K) for u in range(rank, T * K, world)]


def shard_plan(U: int, tiles: int, world: int, min_units_per_rank: int = 16) -> Tuple[int, int]:
    """(ranks over units, ranks over site-pattern blocks) for ``world`` GPUs.  Units shard without any exchange, so they
    are preferred whenever every rank still gets a reasonable number of them; otherwise (one big tree, few samples: the
    config-2 shape) every rank evaluates ALL units over its own contiguous range of 256-pattern tiles and the per-unit
    partial results [ll_data, d/d bl] are summed by one extra all-reduce inside the step."""
    if world == 1 or U >= min_units_per_rank * world or tiles < world:
        return world, 1
    return 1, world


def tile_range(tiles: int, part: int, parts: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``tiles`` 256-pattern blocks: the range owned by ``part`` of ``parts``."""
    return (tiles * part) // parts, (tiles * (part + 1)) // parts


def payload_count(T: int, n: int, m: int) -> int:
    return 3 + 4 * m + T * (n - 1)


def reduce_elbo(unit_tree: torch.Tensor, ll: torch.Tensor, grads: Dict[str, torch.Tensor], T: int, n: int, m: int, K: int,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``vbsky_reduce_elbo``: per-unit ll [U] and gradients (``Ensemble.GRAD_KEYS`` -> [U, ...]) of this rank's units ->
    the all-reduce payload [sum ll / K, sum d/d(origin, clock_rate, R, delta, s, rho) / K, per-tree sum of
    d/d(proportions, root_proportion) / K].  ``unit_tree`` (int32 [U]) must be non-decreasing."""
    dev = ll.device
    if out is None:
        out = torch.empty(payload_count(T, n, m), dtype=torch.float64, device=dev)
    cg = _lib.ParamGrads()
    keep = []
    for k in ("proportions", "root_proportion", "origin", "clock_rate", "R", "delta", "s", "rho"):
        v = grads.get(k)
        if v is not None and v.numel():
            v = v.contiguous()
            keep.append(v)
            setattr(cg, k, v.data_ptr())
        else:
            setattr(cg, k, 0)
    with torch.cuda.device(dev):
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(_lib.lib.vbsky_reduce_elbo(ll.shape[0], T, n, m, C.c_void_p(unit_tree.data_ptr()), C.c_void_p(ll.data_ptr()),
                                              C.byref(cg), 1.0 / K, C.c_void_p(out.data_ptr()), st))
    return out


def allreduce_elbo(vec: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the payload over ranks in place (no-op without an initialised process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def unpack_payload(vec: torch.Tensor, T: int, n: int, m: int) -> Dict[str, torch.Tensor]:
    out = {"elbo": vec[0], "origin": vec[1], "clock_rate": vec[2]}
    for i, k in enumerate(("R", "delta", "s", "rho")):
        out[k] = vec[3 + i * m: 3 + (i + 1) * m]
    out["local"] = vec[3 + 4 * m:].view(T, n - 1)  # [:, :n-2] proportions, [:, n-2] root_proportion
    return out
