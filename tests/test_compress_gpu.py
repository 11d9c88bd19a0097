"""GPU tip encoding + site-pattern compression: bit-exact with np.unique(axis=0,
return_counts=True) on the unpacked partials (fasta.py:539-548) and with the host C path."""
import json
import os

import numpy as np
import pytest
import torch

import oracle.substitution as osub

pytestmark = pytest.mark.gpu


def _np_unique(codes, perm=None):
    part = ((codes[..., None] >> np.arange(4, dtype=np.uint8)) & 1).astype(np.float64)
    if perm is not None:
        part = part[:, perm]
    pat, cnt = np.unique(part, axis=0, return_counts=True)
    return (pat.astype(np.uint8) << np.arange(4, dtype=np.uint8)).sum(-1).astype(np.uint8), cnt


@pytest.mark.parametrize("T,L,n", [(1, 1, 1), (3, 37, 5), (2, 2000, 63), (2, 29903, 200)])
def test_compress_matches_np_unique(cuda, T, L, n):
    from vbsky_b200.tips import compress_patterns, compress_patterns_device

    rng = np.random.default_rng(L + n)
    # few variable columns so that duplicates are common, as in real alignments
    base = rng.choice(np.array([1, 2, 4, 8], dtype=np.uint8), (T, 1, n))
    codes = np.repeat(base, L, axis=1)
    var = rng.random((T, L, n)) < (3.0 / max(n, 3))
    codes[var] = rng.choice(np.array([1, 2, 4, 8, 15, 5, 10], dtype=np.uint8), int(var.sum()))
    perm = np.stack([rng.permutation(n) for _ in range(T)]).astype(np.int32)
    out = compress_patterns_device(torch.as_tensor(codes, device=cuda), torch.as_tensor(perm, device=cuda))
    for t in range(T):
        pat, cnt = out[t]
        want_p, want_c = _np_unique(codes[t], perm[t])
        assert np.array_equal(pat.cpu().numpy(), want_p)
        assert np.array_equal(cnt.cpu().numpy(), want_c)
        hp, hc = compress_patterns(codes[t], perm[t])
        assert np.array_equal(hp, want_p) and np.array_equal(hc, want_c)


def test_encode_and_hcv_fixture(cuda):
    from vbsky_b200.tips import compress_patterns_device, encode_codes_device

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hcv_patterns.json")))
    codes = encode_codes_device(gold["seqs"], cuda)
    want = np.array([osub.encode_partials(s) for s in gold["seqs"]]).transpose(1, 0, 2)
    got = ((codes.cpu().numpy()[..., None] >> np.arange(4, dtype=np.uint8)) & 1).astype(np.float64)
    assert np.array_equal(got, want)
    pat, cnt = compress_patterns_device(codes)
    assert pat.shape == (246, 63) and cnt.cpu().tolist() == gold["counts"]
    assert [int(x) for x in pat[:, 0].cpu()] == gold["first_column_codes"]
    with pytest.raises(KeyError):
        encode_codes_device(["ACgT", "ACGT"], cuda)
