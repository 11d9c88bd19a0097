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
