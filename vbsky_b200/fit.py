"""ELBO over an ensemble and the optimiser loop on top of the CUDA path.

Host-side mirror of ``optim.loss`` (optim.py:35-78) and of ``SeqData.setup_flows`` / ``SeqData.loop``
(fasta.py:362-487).  The variational family is declared the way the reference's scripts declare it -- one entry per
model parameter, either a constant or a mean-field Gaussian pushed through ``pos = Compose(DiagonalAffine, Exp)`` or
``z1 = Compose(DiagonalAffine, ZeroOne)`` (fasta.py:61-65, example.ipynb) -- plus the per-tree local family
``proportions ~ z1(n-2)``, ``root_proportion ~ z1(1)`` (fasta.py:378-384).

    loss_j = -( -sum_flows mean_k log q(x_k) + mean_k [prior(theta_k) + tree prior + data likelihood] )       (optim.py:35-78)

``elbo(trees=[j])`` is exactly ``-loss_j`` of the reference for tree j, with the K draws playing the role of the M = 10
single-sample evaluations that ``step`` averages (fasta.py:403-427; the two are the same number for the same draws).
``elbo()`` sums the likelihood terms and local entropies over all trees at common global draws -- the batched
objective the multi-GPU design evaluates.

The few elementwise maps of the *global* flows stay in PyTorch on the device (the reference keeps them in JAX); the
local flows, the height transform, the tree prior and the pruning likelihood run in the kernels.  Draws can be
injected (``eps=``) so that parity tests can replay the reference's own base draws.
"""
import math
from typing import Callable, Dict, Optional, Sequence, Tuple

import torch

from . import bdsky
from .ensemble import Ensemble
from .flows import local_flow_rows
from .optim import Adagrad

FlowSpec = Tuple[str, object]  # ("pos" | "z1", dim)  or  ("const", value(s))
MODEL_KEYS = ("origin", "origin_start", "clock_rate", "R", "delta", "s", "rho_m")
_LOG_2PI = math.log(2 * math.pi)


def default_global_flows(m: int, origin_start: float, clock_rate: float = 1.10e-3) -> Dict[str, FlowSpec]:
    """The defaults of ``SeqData.setup_flows`` (fasta.py:366-375)."""
    return {"origin": ("pos", 1), "origin_start": ("const", origin_start), "delta": ("const", [36.5] * m), "R": ("pos", m),
            "rho_m": ("const", 0.0), "s": ("const", [0.02] * m), "precision": ("pos", 1), "clock_rate": ("const", clock_rate)}


def _flow_sample(kind: str, p: Dict[str, torch.Tensor], eps: torch.Tensor):
    """Reparameterised sample and its log-density for ``Transform(dim, Compose(DiagonalAffine, Exp | ZeroOne))`` over a
    MeanField base (transform.py:34-65, 130-202): x = g(mu + sigma * eps), log q(x) = log N(eps) - sum(log_sigma) - ldj_g."""
    z = p["mu"] + torch.exp(p["log_sigma"]) * eps
    base = ((eps * eps + _LOG_2PI) / -2).sum(-1) - p["log_sigma"].sum()
    if kind == "pos":
        return torch.exp(z), base - z.sum(-1)
    if kind == "z1":
        sg = torch.sigmoid(z)
        return torch.clamp(sg, 1e-7, 1 - 1e-7), base - (torch.log(sg) + torch.log1p(-sg)).sum(-1)
    raise ValueError(f"unknown flow kind {kind!r}")


class EnsembleELBO:
    """Variational parameters (one flat fp64 leaf ``theta``), sampling, the ELBO and the Adagrad loops.

    ``global_flows``: ordered ``{name: ("pos"|"z1", dim) | ("const", value)}``; must define ``origin``, ``origin_start``,
    ``clock_rate``, ``R``, ``delta``, ``s`` and ``rho_m`` (names as in fasta.py:366-375); any further entry (e.g.
    ``precision``) only feeds ``prior_fn`` and the entropy term.  ``origin_start`` is ONE dataset-level constant
    (``SeqData.earliest``, fasta.py:164,172,368), shared by every tree.  ``prior_fn(samples) -> [K]`` receives
    ``{name: [K, dim]}`` (plus ``rho`` [K, m]) as CUDA fp64 tensors -- the reference's ``_params_prior_loglik``.

    Without ``global_flows`` a minimal family is built from the keyword arguments: origin ~ pos(1), R ~ pos(m), the
    rest constant, ``origin_start`` = the largest sample time of the ensemble.
    """

    def __init__(self, ens: Ensemble, K: int, global_flows: Optional[Dict[str, FlowSpec]] = None,
                 prior_fn: Optional[Callable[[Dict[str, torch.Tensor]], torch.Tensor]] = None, seed: int = 0, *,
                 m: Optional[int] = None, delta: float = 36.5, s: float = 0.02, clock_rate: float = 1.10e-3,
                 origin_start: Optional[float] = None):
        self.ens, self.K = ens, K
        dev = ens.device
        T, n = ens.T, ens.n
        if global_flows is None:
            if m is None:
                raise ValueError("give either global_flows or m")
            if origin_start is None:
                origin_start = max(float(td.sample_times.max()) for td in ens.tds)
            global_flows = {"origin": ("pos", 1), "R": ("pos", m), "origin_start": ("const", origin_start),
                            "delta": ("const", [delta] * m), "s": ("const", [s] * m), "rho_m": ("const", 0.0),
                            "clock_rate": ("const", clock_rate)}
        missing = [k for k in MODEL_KEYS if k not in global_flows]
        if missing:
            raise ValueError(f"global_flows lacks {missing}")
        self.flows = dict(global_flows)
        self._const = {}
        self._shapes = {}
        for name, (kind, arg) in self.flows.items():
            if kind == "const":
                self._const[name] = torch.as_tensor(arg, dtype=torch.float64, device=dev).reshape(-1)
            elif kind in ("pos", "z1"):
                self._shapes[name] = (int(arg),)
            else:
                raise ValueError(f"flow {name}: unknown kind {kind!r}")
        dim = lambda k: self._const[k].numel() if k in self._const else self._shapes[k][0]
        self.m = dim("R")
        if not (dim("delta") == dim("s") == self.m):
            raise ValueError("R, delta and s must have the same number of skyline intervals (bdsky.py:283)")
        for k in ("origin", "origin_start", "clock_rate", "rho_m"):
            if dim(k) != 1:
                raise ValueError(f"{k} must be one-dimensional")
        self.n_global = 2 * sum(sh[0] for sh in self._shapes.values())
        self._shapes["local"] = (T, n - 1)  # local = [proportions (n-2), root_proportion]
        self.unit_tree = torch.arange(T, dtype=torch.int32, device=dev).repeat_interleave(K)
        self.prior_fn = prior_fn
        self.gen = torch.Generator(device=dev)
        self.gen.manual_seed(seed)
        # one flat leaf so that a single fused kernel updates it (the reference initialises with N(0, 0.1), fasta.py:439-446)
        total = 2 * sum(math.prod(sh) for sh in self._shapes.values())
        self.theta = 0.1 * torch.randn(total, dtype=torch.float64, device=dev, generator=self.gen)
        self.theta.requires_grad_(True)

    # ------------------------------------------------------------------------------------------------------------
    @property
    def params(self) -> Dict[str, Dict[str, torch.Tensor]]:
        """Views of the flat parameter block: {flow: {"mu", "log_sigma"}}; per flow ``mu`` then ``log_sigma``."""
        out, off = {}, 0
        for name, sh in self._shapes.items():
            out[name] = {}
            for k in ("mu", "log_sigma"):
                cnt = math.prod(sh)
                out[name][k] = self.theta[off: off + cnt].view(sh)
                off += cnt
        return out

    def leaves(self):
        return [self.theta]

    def set_params(self, values: Dict[str, Dict[str, object]], local: Optional[Sequence[Dict[str, Dict[str, object]]]] = None):
        """Overwrite variational parameters: ``values[flow][mu|log_sigma]``; ``local[t]`` = {"proportions": {...},
        "root_proportion": {...}} in the reference's per-tree layout."""
        n = self.ens.n
        with torch.no_grad():
            par = self.params
            as_t = lambda v: torch.as_tensor(v, dtype=torch.float64, device=self.theta.device).reshape(-1)
            for name, d in values.items():
                for k in ("mu", "log_sigma"):
                    par[name][k].copy_(as_t(d[k]))
            if local is not None:
                for t, lp in enumerate(local):
                    for k in ("mu", "log_sigma"):
                        par["local"][k][t, : n - 2] = as_t(lp["proportions"][k])
                        par["local"][k][t, n - 2] = as_t(lp["root_proportion"][k])[0]

    def sample_eps(self, n_trees: int) -> Dict[str, torch.Tensor]:
        dev, K = self.ens.device, self.K
        rnd = lambda *shape: torch.randn(*shape, dtype=torch.float64, device=dev, generator=self.gen)
        eps = {name: rnd(K, sh[0]) for name, sh in self._shapes.items() if name != "local"}
        eps["local"] = rnd(n_trees * K, self.ens.n - 1)
        return eps

    # ------------------------------------------------------------------------------------------------------------
    def elbo(self, trees: Optional[Sequence[int]] = None, eps: Optional[Dict[str, torch.Tensor]] = None,
             shard: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """ELBO of the whole ensemble, or of the given trees only (``-loss`` of optim.py:35-78 for one tree).
        ``eps``: base draws {flow: [K, dim], "local": [U, n-1]} (units tree-major, K contiguous); drawn if omitted.
        ``shard=(rank, world)``: this rank's share -- the likelihood and local-entropy terms of units u with
        u % world == rank plus 1/world of the global entropy and prior terms -- so that the sum over ranks is the ELBO
        and the sum of ``theta.grad`` over ranks its gradient (every rank must use the same ``eps``)."""
        ens, K, m = self.ens, self.K, self.m
        n, dev = ens.n, ens.device
        par = self.params
        mu_l, ls_l = par["local"]["mu"], par["local"]["log_sigma"]
        if trees is None:
            T, unit_tree = ens.T, self.unit_tree
        else:
            T = len(trees)
            idx = torch.as_tensor(list(trees), dtype=torch.long, device=dev)
            unit_tree = idx.to(torch.int32).repeat_interleave(K)
            mu_l, ls_l = mu_l[idx], ls_l[idx]
        U = T * K
        if eps is None:
            eps = self.sample_eps(T)
        rank, world = (0, 1) if shard is None else shard
        samples: Dict[str, torch.Tensor] = {}
        val = torch.zeros((), dtype=torch.float64, device=dev)
        for name, (kind, _) in self.flows.items():
            if kind == "const":
                samples[name] = self._const[name].reshape(1, -1).expand(K, -1)  # Constant.log_pdf = 0 (distribution.py:122)
                continue
            e = torch.as_tensor(eps[name], dtype=torch.float64, device=dev).reshape(K, -1)
            samples[name], lq = _flow_sample(kind, par[name], e)
            val = val - lq.mean()  # optim.py:60-65
        samples["rho"] = torch.cat([torch.zeros(K, m - 1, dtype=torch.float64, device=dev), samples["rho_m"]], 1)  # optim.py:27-30
        # local family: one row of (mu, log_sigma) per unit; expand's backward is a fixed-order sum over the K draws
        d = n - 1
        e_l = torch.as_tensor(eps["local"], dtype=torch.float64, device=dev).reshape(U, d)
        mu_u, ls_u = mu_l.unsqueeze(1).expand(T, K, d).reshape(U, d), ls_l.unsqueeze(1).expand(T, K, d).reshape(U, d)
        per_unit = lambda v: v.unsqueeze(0).expand(T, K, v.shape[-1]).reshape(U, v.shape[-1])
        if world > 1:
            mine = torch.arange(rank, U, world, device=dev)  # ascending: unit_tree stays grouped by tree
            mu_u, ls_u, e_l, unit_tree = mu_u[mine], ls_u[mine], e_l[mine], unit_tree[mine]
            full = per_unit
            per_unit = lambda v: full(v)[mine]
        x, lq_local = local_flow_rows(mu_u, ls_u, e_l)
        p = {
            "proportions": x[:, : n - 2], "root_proportion": x[:, n - 2],
            "origin": per_unit(samples["origin"])[:, 0], "origin_start": per_unit(samples["origin_start"])[:, 0],
            "clock_rate": per_unit(samples["clock_rate"])[:, 0], "R": per_unit(samples["R"]),
            "delta": per_unit(samples["delta"]), "s": per_unit(samples["s"]), "rho": per_unit(samples["rho"]),
        }
        ll = bdsky.loglik(p, ens, unit_tree, c=(False, True, True), clean=True)
        if self.prior_fn is not None:
            val = val + self.prior_fn(samples).mean()
        if world > 1:
            return val / world + (ll.sum() - lq_local.sum()) / K
        return val + ll.view(T, K).sum(0).mean() - lq_local.view(T, K).mean(1).sum()

    # ------------------------------------------------------------------------------------------------------------
    def fit(self, n_iter: int = 10, step_size: float = 1.0, callback=None, sequential: bool = False, eps_fn=None):
        """The loop of SeqData.loop on -ELBO (fasta.py:387-487): gradients cleaned with nan_to_num (:421,427) and
        jax.example_libraries.optimizers.adagrad(step_size) -- momentum 0.9 on the normalised gradient -- applied by
        a fused kernel (``vbsky_adagrad_update``).

        ``sequential=False`` (default): one update per iteration from the batched ELBO of all trees (what the
        multi-GPU path evaluates); returns the losses [n_iter].  ``sequential=True``: the reference's schedule -- trees
        are visited one at a time, each visit updates the global parameters and that tree's local parameters only,
        with separate optimiser states (fasta.py:448-449, 456-483); returns losses [n_iter][T].
        ``eps_fn(i, j) -> eps`` (j = None in batched mode) injects the base draws of a visit."""
        if not sequential:
            import torch.distributed as dist

            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
            shard = (dist.get_rank(), world) if world > 1 else None
            opt = Adagrad(self.theta.data, step_size=step_size)
            hist = []
            for i in range(n_iter):
                self.theta.grad = None
                # multi-GPU: every rank draws the same eps (same seed), evaluates its units and the gradient of the
                # whole parameter block -- global AND per-tree -- is summed over ranks, so all replicas stay identical
                loss = -self.elbo(eps=None if eps_fn is None else eps_fn(i, None), shard=shard)
                loss.backward()
                if world > 1:
                    red = torch.cat([self.theta.grad, loss.detach().reshape(1)])
                    dist.all_reduce(red)
                    self.theta.grad.copy_(red[:-1])
                    loss = red[-1]
                opt.update(self.theta.grad)
                hist.append(loss.item())
                if callback:
                    callback(i, loss.item())
            return hist
        T, d = self.ens.T, self.ens.n - 1
        G = self.n_global  # the global flows' blocks come first in theta
        blocks = lambda j: (slice(0, G), slice(G + j * d, G + (j + 1) * d), slice(G + (T + j) * d, G + (T + j + 1) * d))
        opt_global = Adagrad(self.theta.data[:G], step_size=step_size)
        opt_local = [[Adagrad(self.theta.data[sl], step_size=step_size) for sl in blocks(j)[1:]] for j in range(T)]
        hist = []
        for i in range(n_iter):
            row = []
            for j in range(T):
                self.theta.grad = None
                loss = -self.elbo(trees=[j], eps=None if eps_fn is None else eps_fn(i, j))
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
