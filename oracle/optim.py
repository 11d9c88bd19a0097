"""TEST INFRASTRUCTURE ONLY -- the optimiser the reference drives in SeqData.loop (fasta.py:388, 421-433):
``jax.example_libraries.optimizers.adagrad(step_size, momentum=0.9)`` on ``nan_to_num`` gradients.  JAX is absent from
this image; the update rule is restated from its published source (parity unpinned by reference outputs)."""
import numpy as np


def adagrad_init(x0):
    return np.array(x0, dtype=np.float64), np.zeros_like(x0, dtype=np.float64), np.zeros_like(x0, dtype=np.float64)


def adagrad_update(g, state, step_size, momentum=0.9):
    x, g_sq, m = state
    g = np.nan_to_num(np.asarray(g, dtype=np.float64))  # fasta.py:421,427
    with np.errstate(divide="ignore", over="ignore"):
        g_sq = g_sq + np.square(g)
        g_sq_inv_sqrt = np.where(g_sq > 0, 1.0 / np.sqrt(g_sq), 0.0)
    m = (1.0 - momentum) * (g * g_sq_inv_sqrt) + momentum * m
    x = x - step_size * m
    return x, g_sq, m


# ---------------------------------------------------------------------------------------------------------------------
# optim.loss (optim.py:35-78) and the M-sample step of SeqData.loop (fasta.py:390-436), restated on torch fp64 with the
# base draws ``eps`` as inputs (JAX's PRNG stream is not reproducible here).
# ---------------------------------------------------------------------------------------------------------------------
def loss(params, flows, td, tip_data, eps, Q, c=((True, True), (True, True, True)), equidistant_intervals=True,
         params_prior_loglik=lambda p: 0.0):
    """optim.py:35-78 for one reparameterised sample per flow (``v.sample(subrng, p, 1)``, :58).

    ``flows``: ordered {name: ("pos" | "z1", dim) or ("const", value)} -- the families fasta.py:366-384 builds
    (pos = Compose(DiagonalAffine, Exp), z1 = Compose(DiagonalAffine, ZeroOne), Constant);
    ``params[name]`` = {"mu", "log_sigma"} torch tensors for the Transform flows; ``eps[name]`` = the [dim] base draw.
    Returns -(elbo1 + elbo2)."""
    import torch

    from . import bdsky as obdsky
    from . import prob as oprob

    c1, c2 = c
    samples = {}
    elbo1 = torch.zeros((), dtype=torch.float64)
    for k, (kind, arg) in flows.items():
        if kind == "const":
            cval = torch.as_tensor(arg, dtype=torch.float64).reshape(-1)
            samples[k] = cval.clone().reshape(1, -1)  # jnp.full((n, dim), c), distribution.py:119-120
            if c1[0]:
                elbo1 = elbo1 - oprob.constant_log_pdf(cval, samples[k][0])
            continue
        p = params[k]
        e = torch.as_tensor(eps[k], dtype=torch.float64).reshape(-1)
        assert e.numel() == arg
        if kind == "pos":
            x = oprob.pos_sample(p, e)
            lq = oprob.pos_log_pdf(p, x)
        elif kind == "z1":
            x = oprob.z1_sample(p, e)
            lq = oprob.z1_log_pdf(p, x)
        else:
            raise ValueError(kind)
        samples[k] = x.reshape(1, -1)
        if c1[0]:
            elbo1 = elbo1 - lq  # optim.py:60-65 (mean over the single sample)
    elbo2 = torch.zeros((), dtype=torch.float64)
    if c1[1]:
        un = obdsky.unpack(samples)  # optim.py:19-32
        s2 = obdsky.loglik({k: v[0] for k, v in un.items()}, td, tip_data, Q, c2, True, equidistant_intervals,
                           params_prior_loglik)
        elbo2 = elbo2 + s2  # mean over the single sample (optim.py:72-75)
    return -(elbo1 + elbo2)


def step_value_and_grad(params, flows, td, tip_data, eps_list, Q, c=((True, True), (True, True, True)),
                        params_prior_loglik=lambda p: 0.0):
    """The body of ``step`` in SeqData.loop (fasta.py:403-427): M single-sample ``value_and_grad(loss)`` evaluations,
    ``nan_to_num`` on each (f, g), averaged over M, ``nan_to_num`` on the averaged gradient.
    ``params``: {flow: {"mu", "log_sigma"}} torch leaves; returns (f, {flow: {"mu": g, "log_sigma": g}})."""
    import torch

    M = len(eps_list)
    leaves = [(k, kk) for k in params for kk in params[k]]
    f_tot = 0.0
    g_tot = {kl: torch.zeros_like(params[kl[0]][kl[1]]) for kl in leaves}
    for eps in eps_list:
        p = {k: {kk: v.detach().clone().requires_grad_(True) for kk, v in d.items()} for k, d in params.items()}
        f = loss(p, flows, td, tip_data, eps, Q, c, True, params_prior_loglik)
        gs = torch.autograd.grad(f, [p[k][kk] for k, kk in leaves], allow_unused=True)
        f_tot = f_tot + float(np.nan_to_num(f.item()))
        for kl, g in zip(leaves, gs):
            if g is not None:
                g_tot[kl] = g_tot[kl] + torch.nan_to_num(g)
    out = {}
    for (k, kk), g in g_tot.items():
        out.setdefault(k, {})[kk] = torch.nan_to_num(g / M)
    return f_tot / M, out
