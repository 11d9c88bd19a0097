"""Substitution models and tip encoding: host-side mirror of
/root/reference/vbsky/substitution.py (SubstitutionModel, HKY, JC69, encode_partials),
backed by the C ABI (vbsky_hky, vbsky_encode_tips)."""
import ctypes as C
from typing import NamedTuple

import numpy as np

from . import _lib


class SubstitutionModel(NamedTuple):
    """(pi, P, d, Pinv) with expm(t) = P diag(exp(t d)) Pinv (substitution.py:51-62)."""

    pi: np.ndarray
    P: np.ndarray
    d: np.ndarray
    Pinv: np.ndarray

    @property
    def Q(self) -> np.ndarray:
        q = np.zeros(16)
        _lib.check(_lib.lib.vbsky_subst_rate_matrix(C.byref(self.c_struct()), _lib.np_ptr(q, C.c_double)))
        return q.reshape(4, 4)

    def expm(self, t: float) -> np.ndarray:
        return self.P @ np.diag(np.exp(t * self.d)) @ self.Pinv

    def c_struct(self) -> "_lib.SubstModel":
        m = _lib.SubstModel()
        m.pi[:] = np.asarray(self.pi, dtype=np.float64).ravel()
        m.P[:] = np.asarray(self.P, dtype=np.float64).ravel()
        m.d[:] = np.asarray(self.d, dtype=np.float64).ravel()
        m.Pinv[:] = np.asarray(self.Pinv, dtype=np.float64).ravel()
        return m


def HKY(kappa: float = 1) -> SubstitutionModel:
    """HKY(kappa) with uniform base frequencies, normalised as the reference does
    (substitution.py:73-87): transversion 1/(2+kappa), transition kappa/(2+kappa), diagonal -1."""
    m = _lib.SubstModel()
    _lib.check(_lib.lib.vbsky_hky(float(kappa), C.byref(m)))
    return SubstitutionModel(
        np.array(m.pi[:]), np.array(m.P[:]).reshape(4, 4), np.array(m.d[:]), np.array(m.Pinv[:]).reshape(4, 4)
    )


def HKY_dkappa(kappa: float) -> np.ndarray:
    """d(d)/d(kappa) of ``HKY(kappa)``'s spectrum [4]; dQ/dkappa = P diag(.) Pinv since the eigenvectors are fixed."""
    dd = np.zeros(4)
    _lib.check(_lib.lib.vbsky_hky_dkappa(float(kappa), _lib.np_ptr(dd, C.c_double)))
    return dd


def JC69() -> SubstitutionModel:
    return HKY(1.0)


_PAIRS = ((0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3))  # AC, AG, AT, CG, CT, GT


def gtr_rate_matrix(rates, pi) -> np.ndarray:
    """Q of the general time-reversible model: Q_ij = r_ij pi_j (i != j), rows summing to zero; ``rates`` in the order
    AC, AG, AT, CG, CT, GT (HKY with base frequencies is rates = (1, kappa, 1, 1, kappa, 1)).  Not normalised."""
    rates, pi = np.asarray(rates, dtype=np.float64), np.asarray(pi, dtype=np.float64)
    Q = np.zeros((4, 4))
    for r, (i, j) in zip(rates, _PAIRS):
        Q[i, j] = r * pi[j]
        Q[j, i] = r * pi[i]
    Q -= np.diag(Q.sum(1))
    return Q


def GTR(rates, pi) -> SubstitutionModel:
    """Reversible model with arbitrary exchangeabilities and base frequencies as a (pi, P, d, Pinv) model: the rate matrix
    is similar to the symmetric D^1/2 Q D^-1/2 (D = diag(pi)), whose eigh gives P = D^-1/2 V, Pinv = V^T D^1/2.
    The reference only constructs uniform-frequency HKY (substitution.py:73-92); the kernels take any such model
    (dense-matrix path)."""
    pi = np.asarray(pi, dtype=np.float64)
    Q = gtr_rate_matrix(rates, pi)
    sq = np.sqrt(pi)
    S = (sq[:, None] * Q) / sq[None, :]
    d, V = np.linalg.eigh((S + S.T) / 2)
    P, Pinv = V / sq[:, None], V.T * sq[None, :]
    assert np.allclose(P @ np.diag(d) @ Pinv, Q, atol=1e-12)
    return SubstitutionModel(pi, P, d, Pinv)


def gtr_directions(rates, pi):
    """The derivative directions of ``gtr_rate_matrix``: a list of (name, dQ/dtheta [4, 4], dpi/dtheta [4]) for the six
    exchangeabilities and for the three free base-frequency coordinates (pi_i, i = A, C, G, with pi_T = 1 - sum)."""
    rates, pi = np.asarray(rates, dtype=np.float64), np.asarray(pi, dtype=np.float64)
    out = []
    for k, (i, j) in enumerate(_PAIRS):
        dQ = np.zeros((4, 4))
        dQ[i, j], dQ[j, i] = pi[j], pi[i]
        dQ -= np.diag(dQ.sum(1))
        out.append((f"rate_{'ACGT'[i]}{'ACGT'[j]}", dQ, np.zeros(4)))
    R = np.zeros((4, 4))
    for r, (i, j) in zip(rates, _PAIRS):
        R[i, j] = R[j, i] = r
    for i in range(3):
        dpi = np.zeros(4)
        dpi[i], dpi[3] = 1.0, -1.0
        dQ = R * dpi[None, :]
        dQ -= np.diag(dQ.sum(1))
        out.append((f"pi_{'ACGT'[i]}", dQ, dpi))
    return out


def encode_codes(seqs) -> np.ndarray:
    """n equal-length sequences -> uint8 [L, n] state masks (bit0=A, bit1=C, bit2=G, bit3=T) following
    the BEAST ambiguity table of substitution.py:10-35.  Unknown characters raise KeyError."""
    seqs = [s if isinstance(s, str) else "".join(s) for s in seqs]
    n = len(seqs)
    L = len(seqs[0]) if n else 0
    if any(len(s) != L for s in seqs):
        raise ValueError("sequences must have equal length")
    buf = "".join(seqs).encode("latin-1")
    codes = np.zeros((L, n), dtype=np.uint8)
    _lib.check(_lib.lib.vbsky_encode_tips(buf, n, L, _lib.np_ptr(codes, C.c_uint8)))
    return codes


def codes_to_partials(codes: np.ndarray) -> np.ndarray:
    """uint8 masks [...] -> float64 0/1 partials [..., 4] (the reference's dense representation)."""
    codes = np.asarray(codes, dtype=np.uint8)
    return ((codes[..., None] >> np.arange(4, dtype=np.uint8)) & 1).astype(np.float64)


def partials_to_codes(partials: np.ndarray) -> np.ndarray:
    """0/1 partials [..., 4] -> uint8 masks [...]; anything that is not exactly 0 or 1 is rejected."""
    partials = np.asarray(partials)
    if partials.shape[-1] != 4 or not np.isin(partials, (0, 1)).all():
        raise ValueError("tip partials must be 0/1 indicator rows of length 4")
    return (partials.astype(np.uint8) << np.arange(4, dtype=np.uint8)).sum(-1).astype(np.uint8)


def encode_partials(col: str) -> np.ndarray:
    """[n,4] tip partials of one sequence string (substitution.py:38-48)."""
    return codes_to_partials(encode_codes([col])[:, 0])
