"""Batched ELBO over an ensemble and a minimal optimiser loop on top of the CUDA path.

Host-side mirror of ``optim.loss`` (optim.py:35-78) and of the Adagrad loop of ``SeqData.loop``
(fasta.py:387-487) for the *batched* objective the multi-GPU design evaluates: all trees at the same
global parameters, K Monte-Carlo samples each (the reference visits trees one at a time).

    ELBO = mean_k [ sum_t loglik(theta_k, local_{t,k}) + log prior(theta_k) ]
           - mean_k log q_global(theta_k) - sum_t mean_k log q_t(local_{t,k})

Global flows are the reference's defaults (fasta.py:366-375): origin ~ pos(1), R ~ pos(m) with
pos = Compose(DiagonalAffine, Exp); delta, s, rho_m, clock_rate constants.  Those few elementwise maps stay in
PyTorch (the reference keeps them in JAX); the local flows, heights, tree prior and pruning run in the kernels.
"""
import math
from typing import Callable, Dict, Optional, Sequence

import torch

from . import bdsky
from .ensemble import Ensemble
from .flows import local_flow
from .optim import Adagrad


def _pos_sample(p: Dict[str, torch.Tensor], eps: torch.Tensor):
    """Transform(dim, Compose(DiagonalAffine, Exp)) over MeanField (transform.py:130-166): x = exp(mu + sigma*eps),
    log q(x) = log N(eps) - sum(log_sigma) - sum(mu + sigma*eps)."""
    z = p["mu"] + torch.exp(p["log_sigma"]) * eps
    logq = (-0.5 * eps**2 - 0.5 * math.log(2 * math.pi)).sum(-1) - p["log_sigma"].sum() - z.sum(-1)
    return torch.exp(z), logq


class EnsembleELBO:
    def __init__(self, ens: Ensemble, K: int, m: int, delta: float = 36.5, s: float = 0.02, clock_rate: float = 1.12e-3,
                 prior_fn: Optional[Callable[[Dict[str, torch.Tensor]], torch.Tensor]] = None, seed: int = 0):
        self.ens, self.K, self.m = ens, K, m
        dev = ens.device
        T, n = ens.T, ens.n
        self.unit_tree = torch.arange(T, dtype=torch.int32, device=dev).repeat_interleave(K)
        self.origin_start = torch.tensor([float(td.sample_times.max()) for td in ens.tds], dtype=torch.float64, device=dev)
        self.const = dict(delta=delta, s=s, clock_rate=clock_rate)
        self.prior_fn = prior_fn
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        # variational parameters live in one flat leaf so that a single fused kernel updates them
        # (the reference initialises them with N(0, 0.1); fasta.py:439-446)
        self._shapes = {"origin": (1,), "R": (m,), "local": (T, n - 1)}  # local = [proportions (n-2), root_proportion]
        total = 2 * sum(math.prod(sh) for sh in self._shapes.values())
        self.theta = 0.1 * torch.randn(total, dtype=torch.float64, device=dev, generator=self.gen)
        self.theta.requires_grad_(True)

    @property
    def params(self) -> Dict[str, Dict[str, torch.Tensor]]:
        """Views of the flat parameter block: {flow: {"mu", "log_sigma"}}."""
        out, off = {}, 0
        for name, sh in self._shapes.items():
            out[name] = {}
            for k in ("mu", "log_sigma"):
                cnt = math.prod(sh)
                out[name][k] = self.theta[off : off + cnt].view(sh)
                off += cnt
        return out

    def leaves(self):
        return [self.theta]

    def elbo(self, trees: Optional[Sequence[int]] = None) -> torch.Tensor:
        """ELBO of the whole ensemble, or of the given trees only (the reference's per-tree ``_loss``, optim.py:35-78)."""
        ens, K, m = self.ens, self.K, self.m
        n, dev = ens.n, ens.device
        if trees is None:
            T, unit_tree = ens.T, self.unit_tree
        else:
            T = len(trees)
            unit_tree = torch.as_tensor(list(trees), dtype=torch.int32, device=dev).repeat_interleave(K)
        U = T * K
        rnd = lambda *shape: torch.randn(*shape, dtype=torch.float64, device=dev, generator=self.gen)
        par = self.params
        origin, lq_o = _pos_sample(par["origin"], rnd(K, 1))
        R, lq_r = _pos_sample(par["R"], rnd(K, m))
        x, lq_local = local_flow(par["local"]["mu"], par["local"]["log_sigma"], rnd(U, n - 1), unit_tree)
        ks = torch.arange(K, device=dev).repeat(T)  # sample index of each unit
        full = lambda v: torch.full((U, m), v, dtype=torch.float64, device=dev)
        p = {
            "proportions": x[:, : n - 2], "root_proportion": x[:, n - 2],
            "origin": origin[ks, 0], "origin_start": self.origin_start[unit_tree.long()],
            "clock_rate": torch.full((U,), self.const["clock_rate"], dtype=torch.float64, device=dev),
            "R": R[ks], "delta": full(self.const["delta"]), "s": full(self.const["s"]),
            "rho": torch.zeros(U, m, dtype=torch.float64, device=dev),
        }
        ll = bdsky.loglik(p, ens, unit_tree, c=(False, True, True))
        val = ll.sum() / K - lq_o.mean() - lq_r.mean() - lq_local.sum() / K
        if self.prior_fn is not None:
            val = val + self.prior_fn({"origin": origin, "R": R}).mean()
        return val

    def fit(self, n_iter: int = 10, step_size: float = 1.0, callback=None, sequential: bool = False):
        """The loop of SeqData.loop on -ELBO (fasta.py:387-487): gradients cleaned with nan_to_num (:421,427) and
        jax.example_libraries.optimizers.adagrad(step_size) -- momentum 0.9 on the normalised gradient -- applied by
        a fused kernel (``vbsky_adagrad_update``).

        ``sequential=False`` (default): one update per iteration from the batched ELBO of all trees (what the
        multi-GPU path evaluates); returns the losses [n_iter].  ``sequential=True``: the reference's schedule -- trees
        are visited one at a time, each visit updates the global parameters and that tree's local parameters only,
        with separate optimiser states (fasta.py:448-449, 456-483); returns losses [n_iter][T]."""
        if not sequential:
            opt = Adagrad(self.theta.data, step_size=step_size)
            hist = []
            for i in range(n_iter):
                self.theta.grad = None
                loss = -self.elbo()
                loss.backward()
                opt.update(self.theta.grad)
                hist.append(loss.item())
                if callback:
                    callback(i, loss.item())
            return hist
        T, d = self.ens.T, self.ens.n - 1
        G = 2 * (1 + self.m)  # origin and R blocks come first in theta
        blocks = lambda j: (slice(0, G), slice(G + j * d, G + (j + 1) * d), slice(G + (T + j) * d, G + (T + j + 1) * d))
        opt_global = Adagrad(self.theta.data[:G], step_size=step_size)
        opt_local = [[Adagrad(self.theta.data[sl], step_size=step_size) for sl in blocks(j)[1:]] for j in range(T)]
        hist = []
        for i in range(n_iter):
            row = []
            for j in range(T):
                self.theta.grad = None
                loss = -self.elbo(trees=[j])
                loss.backward()
                g = self.theta.grad
                sl = blocks(j)
                opt_global.update(g[sl[0]])
                opt_local[j][0].update(g[sl[1]])
                opt_local[j][1].update(g[sl[2]])
                row.append(loss.item())
                if callback:
                    callback(i, row[-1])
            hist.append(row)
        return hist
