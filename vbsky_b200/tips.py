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


def compress_patterns_device(codes, perm=None):
    """Device version of ``compress_patterns`` for a batch of alignments: ``codes`` CUDA uint8 tensor
    [T, L, n] of state masks (or [L, n]), optional ``perm`` CUDA int32 [T, n].  Returns a list of
    (patterns uint8 [S_t, n], counts int64 [S_t]) CUDA tensors, bit-identical to the host path."""
    import torch

    single = codes.dim() == 2
    if single:
        codes = codes[None]
        perm = None if perm is None else perm[None]
    codes = codes.contiguous()
    T, L, n = codes.shape
    dev = codes.device
    patterns = torch.empty_like(codes)
    counts = torch.empty((T, L), dtype=torch.int64, device=dev)
    S = torch.empty(T, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nbytes = _lib.lib.vbsky_compress_workspace_bytes(T, L)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        pp = C.c_void_p(0 if perm is None else perm.contiguous().data_ptr())
        _lib.check(_lib.lib.vbsky_compress_patterns_device(
            T, L, n, C.c_void_p(codes.data_ptr()), pp, C.c_void_p(patterns.data_ptr()), C.c_void_p(counts.data_ptr()),
            C.c_void_p(S.data_ptr()), C.c_void_p(ws.data_ptr()), nbytes, st))
    Sh = S.cpu().tolist()
    out = [(patterns[t, : Sh[t]], counts[t, : Sh[t]]) for t in range(T)]
    return out[0] if single else out


def encode_codes_device(seqs, device):
    """Device version of ``encode_codes``: n equal-length sequences -> CUDA uint8 [L, n] masks."""
    import torch

    seqs = [s if isinstance(s, str) else "".join(s) for s in seqs]
    n, L = len(seqs), len(seqs[0])
    if any(len(s) != L for s in seqs):
        raise ValueError("sequences must have equal length")
    raw = torch.frombuffer(bytearray("".join(seqs).encode("latin-1")), dtype=torch.uint8).to(device)
    codes = torch.empty((L, n), dtype=torch.uint8, device=device)
    bad = C.c_int64(-1)
    with torch.cuda.device(device):
        st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _lib.check(_lib.lib.vbsky_encode_tips_device(C.c_void_p(raw.data_ptr()), n, L, C.c_void_p(codes.data_ptr()), C.byref(bad), st))
    return codes
