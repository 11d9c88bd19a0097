"""Tip-data preparation: site-pattern compression and padding, bit-exact with the reference's
``_process_tips`` (fasta.py:539-548) and ``pad_tips`` (fasta.py:325-331), on packed 4-bit codes."""
import ctypes as C
from typing import List, NamedTuple, Optional, Sequence

import numpy as np

from . import _lib
from .substitution import codes_to_partials, encode_codes
from .util import TipData


class PackedTips(NamedTuple):
    """patterns uint8 [S, n] state masks (tips in node order), counts float64 [S]."""

    patterns: np.ndarray
    counts: np.ndarray

    def to_tip_data(self) -> TipData:
        return TipData(codes_to_partials(self.patterns), self.counts.copy())


def compress_patterns(codes: np.ndarray, perm: Optional[Sequence[int]] = None):
    """codes uint8 [L, n] -> (patterns [S, n], counts int64 [S]) == np.unique(partials[:, perm],
    axis=0, return_counts=True) of fasta.py:545-547 (same row order, same counts)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    L, n = codes.shape
    patterns = np.zeros((max(L, 1), n), dtype=np.uint8)
    counts = np.zeros(max(L, 1), dtype=np.int64)
    S = C.c_int64(0)
    pp = None
    if perm is not None:
        perm = np.ascontiguousarray(perm, dtype=np.int64)
        pp = _lib.np_ptr(perm, C.c_int64)
    _lib.check(
        _lib.lib.vbsky_compress_patterns(
            _lib.np_ptr(codes, C.c_uint8), L, n, pp, _lib.np_ptr(patterns, C.c_uint8),
            _lib.np_ptr(counts, C.c_int64), C.byref(S),
        )
    )
    return patterns[: S.value].copy(), counts[: S.value].copy()


def process_tips(sids: Sequence, seqs: Sequence[str], node_mapping: dict) -> PackedTips:
    """fasta.py:539-548: encode, order tips by their node index, compress site patterns."""
    codes = encode_codes(seqs)
    perm = np.argsort([node_mapping[s] for s in sids])
    patterns, counts = compress_patterns(codes, perm)
    return PackedTips(patterns, counts.astype(np.float64))


def pad_tips(tips: List[PackedTips]) -> List[PackedTips]:
    """fasta.py:325-331: pad every tree to the ensemble's largest pattern count with fully ambiguous
    patterns (all-ones partials == mask 15) of count 0."""
    smax = max(t.patterns.shape[0] for t in tips)
    out = []
    for t in tips:
        pad = smax - t.patterns.shape[0]
        patterns = np.concatenate([t.patterns, np.full((pad, t.patterns.shape[1]), 15, dtype=np.uint8)])
        counts = np.concatenate([t.counts, np.zeros(pad)])
        out.append(PackedTips(patterns, counts))
    return out
