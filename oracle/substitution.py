"""Oracle restatement of /root/reference/vbsky/substitution.py (test infrastructure).

Follows substitution.py:10-48 (BEAST ambiguity table, encode_partials) and
substitution.py:51-92 (SubstitutionModel, HKY, JC69).  ``msprime.HKY`` is absent from
this image; its published construction is restated in ``_msprime_hky`` below.
"""
from typing import NamedTuple

import numpy as np

# substitution.py:10-30 -- beast.evolution.datatype.Nucleotide state sets.
_BEAST_STATES = {
    "A": (0,), "C": (1,), "G": (2,), "T": (3,), "U": (3,),
    "R": (0, 2), "Y": (1, 3), "M": (0, 1), "W": (0, 3), "S": (1, 2), "K": (2, 3),
    "B": (1, 2, 3), "D": (0, 2, 3), "H": (0, 1, 3), "V": (0, 1, 2),
    "N": (0, 1, 2, 3), "X": (0, 1, 2, 3), "-": (0, 1, 2, 3), "?": (0, 1, 2, 3),
}

# substitution.py:32-35 -- PARTIAL_DICT: character -> {0,1}^4 indicator (float64).
PARTIAL_DICT = {ch: np.eye(4)[list(st)].sum(0) for ch, st in _BEAST_STATES.items()}


def encode_partials(col: str) -> np.ndarray:
    """substitution.py:38-48 -- [n,4] tip partials for one sequence/column string.

    Upper-case only: a lower-case character raises KeyError exactly as the reference does.
    """
    return np.array(list(map(PARTIAL_DICT.__getitem__, col)))


class SubstitutionModel(NamedTuple):
    """substitution.py:51-70."""

    pi: np.ndarray
    P: np.ndarray
    d: np.ndarray
    Pinv: np.ndarray

    @property
    def Q(self):
        # substitution.py:57-59
        return self.P @ (self.d[:, None] * self.Pinv)

    def expm(self, t):
        # substitution.py:61-62
        return self.P @ np.diag(np.exp(t * self.d)) @ self.Pinv

    def expm_action(self, t, p, right: bool = True):
        # substitution.py:64-70, incl. the clip to [1e-40, 1].
        # p may carry leading batch axes ([..., 4]); the contraction is over the last axis.
        p = np.asarray(p, dtype=np.float64)
        e = np.exp(t * self.d)
        if right:
            ret = (e * (p @ self.Pinv.T)) @ self.P.T
        else:
            ret = ((p @ self.P) * e) @ self.Pinv
        return np.clip(ret, 1e-40, 1.0)


def _msprime_hky(kappa: float):
    """Published construction of ``msprime.HKY(kappa)`` (msprime >= 1.0, mutations.py):

    alleles ACGT, uniform equilibrium frequencies; off-diagonal entry (i,j) is
    pi_j * kappa for transitions (A<->G, C<->T) and pi_j otherwise; the matrix is divided
    by the largest off-diagonal row sum and the diagonal set so every row sums to 1.
    """
    pi = np.array([0.25, 0.25, 0.25, 0.25])
    M = np.full((4, 4), 1.0)
    for i, j in ((0, 2), (1, 3), (2, 0), (3, 1)):
        M[i, j] = kappa
    M = M * pi[None, :]
    np.fill_diagonal(M, 0.0)
    row_sums = M.sum(1)
    M = M / row_sums.max()
    row_sums = row_sums / row_sums.max()
    np.fill_diagonal(M, 1.0 - row_sums)
    return M, pi


def HKY(kappa: float = 1) -> SubstitutionModel:
    """substitution.py:73-87."""
    Q, pi = _msprime_hky(kappa)
    Q = Q - np.diag(Q.sum(1))
    d, P = np.linalg.eigh(Q)
    assert np.allclose(P @ np.diag(d) @ P.T, Q)
    ret = SubstitutionModel(pi, P, d, P.T)
    assert np.allclose(ret.Q, Q)
    return ret


def JC69() -> SubstitutionModel:
    """substitution.py:90-92."""
    return HKY(1.0)
