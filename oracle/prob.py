"""Oracle restatement (torch fp64; test infrastructure) of the reference's *local* variational family
-- Transform(dim, Compose(DiagonalAffine, ZeroOne)) over a MeanField base (fasta.py:63-65,378-384;
prob/transform.py:34-65,130-202,320-438; prob/distribution.py:52-65) -- and of the entropy term of
optim.loss (optim.py:54-65).  The base noise ``eps`` is an input: JAX's PRNG stream is not reproducible here."""
import math

import torch


def diagonal_affine_direct(params, x):  # transform.py:141-142
    return torch.exp(params["log_sigma"]) * x + params["mu"]


def diagonal_affine_inverse(params, y):  # transform.py:147-148
    return (y - params["mu"]) * torch.exp(-params["log_sigma"])


def zero_one_direct(x):  # transform.py:194-195
    return torch.clamp(torch.sigmoid(x), 1e-7, 1 - 1e-7)


def zero_one_inverse(y):  # transform.py:200-202 (jax logit = log(x / (1 - x)))
    yc = torch.clamp(y, 1e-7, 1 - 1e-7)
    return torch.log(yc / (1 - yc))


def zero_one_log_det_jac(x):  # transform.py:197-198
    s = torch.sigmoid(x)
    return (torch.log(s) + torch.log1p(-s)).sum(-1)


def z1_sample(params, eps):
    """TransformedDistribution.sample (transform.py:52-54) for z1 = Compose(DiagonalAffine, ZeroOne)."""
    return zero_one_direct(diagonal_affine_direct(params, eps))


def z1_log_pdf(params, x):
    """TransformedDistribution.log_pdf (transform.py:56-63): f_X(f^-1(y)) / |Df(x)| with the composition's
    inverse / log-det-jacobian (transform.py:350-372) and the MeanField base (distribution.py:63-65)."""
    z = zero_one_inverse(x)
    x0 = diagonal_affine_inverse(params, z)
    base = (-0.5 * x0 * x0 - 0.5 * math.log(2 * math.pi)).sum(-1)
    ldj = params["log_sigma"].sum() + zero_one_log_det_jac(diagonal_affine_direct(params, x0))
    return base - ldj


# ---- Transform(dim, Compose(DiagonalAffine, Exp)) = the reference's ``pos`` family (fasta.py:61, transform.py:152-166)
def norm_logpdf(x):
    """jax.scipy.stats.norm.logpdf with loc 0, scale 1: -(log(2 pi) + x^2) / 2."""
    return (math.log(2 * math.pi) + x * x) / -2


def pos_sample(params, eps):
    return torch.exp(diagonal_affine_direct(params, eps))  # transform.py:158-159


def pos_log_pdf(params, x):
    """TransformedDistribution.log_pdf (transform.py:56-63) for Compose(DiagonalAffine, Exp)."""
    x0 = diagonal_affine_inverse(params, torch.log(x))  # transform.py:164-166, 147-148
    base = norm_logpdf(x0).sum(-1)  # distribution.py:63-65
    ldj = params["log_sigma"].sum() + diagonal_affine_direct(params, x0).sum(-1)  # transform.py:144-145, 161-162
    return base - ldj


def constant_log_pdf(c, x):
    """Constant.log_pdf (distribution.py:122-123)."""
    close = (x - c).abs() <= 1e-8 + 1e-5 * c.abs()
    return torch.where(close, torch.zeros_like(x), torch.full_like(x, -math.inf)).sum()
