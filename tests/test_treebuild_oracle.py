"""CPU checks of the starting-tree oracle (oracle/treebuild.py) against independent computations and hand-traced
cases.  snp-dists and Biopython are absent from the reference tree and this image (parity unpinned); these tests pin
the restated rules the GPU path is compared with."""
import numpy as np
import pytest

import oracle.treebuild as otb


def test_snp_dists_rules():
    # differences count only where both characters are in AGTC after upper-casing
    seqs = ["AACGTN-A", "AaCTTACA", "TRCGAN?C"]
    d = otb.snp_dists(seqs)
    #            cols: 0 1 2 3 4 5 6 7
    # 0 vs 1:          = = = x = i i =   -> 1
    # 0 vs 2:          x i = = x i i x   -> 3
    # 1 vs 2:          x i = x x i i x   -> 4
    assert d.tolist() == [[0, 1, 3], [1, 0, 4], [3, 4, 0]]


@pytest.mark.parametrize("single_theta", [False, True])
def test_distance_matrix_regression_matches_lstsq(single_theta):
    rng = np.random.default_rng(3)
    n = 12
    times = np.round(rng.uniform(0, 2, n), 1)  # repeated time classes
    r, c = np.tril_indices(n, -1)
    y = rng.integers(0, 40, len(r))
    mat, omega = otb.get_distance_matrix(y, times, single_theta)
    dt = np.abs(times[r] - times[c])
    if single_theta:
        X = np.column_stack([np.ones(len(r)), dt])
    else:
        later = np.maximum(times[r], times[c])
        classes = np.unique(times)
        X = np.column_stack([(later[:, None] == classes[None, :]).astype(float), dt])
    beta = np.linalg.lstsq(X, y.astype(float), rcond=None)[0]
    assert omega == pytest.approx(beta[-1], rel=1e-9, abs=1e-12)
    assert mat[r, c] == pytest.approx(y + omega * (times[r] + times[c] - 2 * times.min()))
    assert np.all(np.triu(mat) == 0)


def test_distance_matrix_equal_times_gives_zero_omega():
    y = np.array([3, 5, 4])
    mat, omega = otb.get_distance_matrix(y, np.zeros(3))
    assert omega == 0.0
    assert mat[np.tril_indices(3, -1)].tolist() == [3.0, 5.0, 4.0]


def test_upgma_all_ties_takes_last_scan_hit():
    # every distance equal: Biopython's `>=` scan keeps the last (i, j) of "for i: for j < i"
    mat = np.tril(np.ones((4, 4)), -1)
    merges, heights, tree = otb.upgma(mat, "abcd")
    assert merges.tolist() == [[3, 2], [4, 1], [5, 0]]
    assert heights.tolist() == [0.5, 0.5, 0.5]
    assert tree == ((("d", "c"), "b"), "a")
    assert otb.nested_to_newick(tree) == "(((d,c),b),a);"


def test_upgma_textbook_average():
    # a, b close; c, d close; the new row is the plain average of the merged rows
    D = np.array([[0, 2, 8, 10], [2, 0, 6, 12], [8, 6, 0, 4], [10, 12, 4, 0]], dtype=float)
    merges, heights, tree = otb.upgma(D, "abcd")
    # round 1: min 2 at (1, 0) -> Inner1 = (b, a) at position 0; row: c -> (6+8)/2 = 7, d -> (12+10)/2 = 11
    # round 2: matrix [I1, c, d]: d(c, I1) = 7, d(d, I1) = 11, d(d, c) = 4 -> Inner2 = (d, c) at position 1
    # round 3: d(I2, I1) = (11 + 7)/2 = 9
    assert merges.tolist() == [[1, 0], [3, 2], [5, 4]]
    assert heights.tolist() == [1.0, 2.0, 4.5]
    assert tree == (("d", "c"), ("b", "a"))


def test_upgma_is_a_valid_monotone_clustering():
    rng = np.random.default_rng(0)
    n = 40
    x = rng.random((n, 3))
    D = np.linalg.norm(x[:, None] - x[None], axis=-1)
    merges, heights, _ = otb.upgma(D)
    assert sorted(merges.reshape(-1).tolist()) == list(range(2 * n - 2))  # every clade is used exactly once
    assert all(max(m) < n + k for k, m in enumerate(merges))
    assert np.all(np.diff(heights) >= -1e-12)


@pytest.mark.parametrize("kind", ["float", "int", "ties"])
def test_vectorised_upgma_equals_literal_loop(kind):
    rng = np.random.default_rng(5)
    n = 37
    if kind == "float":
        x = rng.random((n, 3))
        D = np.linalg.norm(x[:, None] - x[None], axis=-1)
    elif kind == "int":
        D = rng.integers(0, 5, (n, n)).astype(float)
    else:
        D = np.ones((n, n))
    D = np.tril(D, -1)
    m1, h1, _ = otb.upgma(D)
    m2, h2 = otb.upgma_vectorised(D)
    assert np.array_equal(m1, m2) and np.array_equal(h1, h2)
