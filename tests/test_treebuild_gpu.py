"""Starting trees on the device against the restated snp-dists / get_distance_matrix /
Biopython-UPGMA chain of oracle/treebuild.py (fasta.py:255-284, upgma.py:25-62): integer distances and merge
lists bit-exact, omega and the corrected distances to 1e-9, and the resulting TreeData arrays identical."""
import json
import os

import numpy as np
import pytest
import torch

import oracle.treebuild as otb
import oracle.tree_data as otd

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _random_alignment(rng, N, L, p_var=0.05):
    root = rng.choice(list("ACGT"), L)
    seqs = np.repeat(root[None], N, axis=0)
    var = rng.random((N, L)) < p_var
    seqs[var] = rng.choice(list("ACGTN-RYacgtn?"), int(var.sum()))
    return ["".join(s) for s in seqs]


@pytest.mark.parametrize("N,L,T,n", [(2, 1, 1, 2), (9, 33, 3, 5), (40, 411, 4, 17), (120, 2999, 3, 64)])
def test_snp_dists_bit_exact(cuda, N, L, T, n):
    from vbsky_b200.treebuild import PackedAlignment, snp_dists

    rng = np.random.default_rng(N + L)
    seqs = _random_alignment(rng, N, L)
    rows = np.stack([rng.choice(N, n, replace=False) for _ in range(T)])
    aln = PackedAlignment(seqs, cuda)
    got = snp_dists(aln, rows).cpu().numpy()
    for t in range(T):
        want = otb.snp_dists([seqs[i] for i in rows[t]])
        assert np.array_equal(got[t], want)


def test_snp_dists_hcv_fixture(cuda):
    from vbsky_b200.treebuild import PackedAlignment, snp_dists

    with open(os.path.join(GOLD, "hcv_patterns.json")) as f:
        fx = json.load(f)
    seqs = fx["seqs"] if "seqs" in fx else None
    if seqs is None:
        pytest.skip("fixture holds no raw sequences")
    aln = PackedAlignment(seqs, cuda)
    got = snp_dists(aln, np.arange(len(seqs))[None]).cpu().numpy()[0]
    assert np.array_equal(got, otb.snp_dists(seqs))


def _times(rng, kind, n):
    if kind == "distinct":
        return rng.uniform(0, 3, n)
    if kind == "classes":
        return np.round(rng.uniform(0, 1, n), 1)
    if kind == "equal":
        return np.full(n, 0.7)
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["distinct", "classes", "equal"])
@pytest.mark.parametrize("single_theta", [False, True])
def test_supgma_distances(cuda, kind, single_theta):
    from vbsky_b200.treebuild import supgma_distances

    rng = np.random.default_rng(11)
    T, n = 3, 23
    r, c = np.tril_indices(n, -1)
    dist = np.zeros((T, n, n), dtype=np.int32)
    times = np.stack([_times(rng, kind, n) for _ in range(T)])
    for t in range(T):
        y = rng.integers(0, 60, len(r)) + (30 * np.abs(times[t][r] - times[t][c])).astype(int)
        dist[t][r, c] = y
        dist[t][c, r] = y
    D, omega = supgma_distances(torch.as_tensor(dist, device=cuda), times, single_theta)
    D, omega = D.cpu().numpy(), omega.cpu().numpy()
    for t in range(T):
        mat, om = otb.get_distance_matrix(dist[t][r, c], times[t], single_theta)
        assert omega[t] == pytest.approx(om, rel=1e-9, abs=1e-12)
        assert np.allclose(D[t][r, c], mat[r, c], rtol=1e-9, atol=1e-9)
        assert np.array_equal(D[t], D[t].T) and np.all(np.diag(D[t]) == 0)


@pytest.mark.parametrize("n,kind", [(2, "float"), (3, "ties"), (17, "float"), (40, "ties"), (64, "int"), (200, "float")])
def test_upgma_bit_exact(cuda, n, kind):
    from vbsky_b200.treebuild import upgma

    rng = np.random.default_rng(n)
    T = 3
    Ds = []
    for t in range(T):
        if kind == "float":
            x = rng.random((n, 4))
            D = np.linalg.norm(x[:, None] - x[None], axis=-1)
        elif kind == "int":  # integer SNP distances with equal sample times: many exact ties
            D = rng.integers(0, 6, (n, n)).astype(float)
        else:
            D = np.ones((n, n))
        D = np.tril(D, -1)
        Ds.append(D + D.T)
    merges, heights = upgma(torch.as_tensor(np.stack(Ds), device=cuda))
    merges, heights = merges.cpu().numpy(), heights.cpu().numpy()
    for t in range(T):
        m, h, _ = otb.upgma(Ds[t])
        assert np.array_equal(merges[t], m)
        assert np.array_equal(heights[t], h)


def test_upgma_beyond_one_position_per_thread(cuda):
    """n > 1024 tips (several live positions per thread) with integer distances, i.e. many exact ties."""
    from vbsky_b200.treebuild import upgma

    rng = np.random.default_rng(3)
    n = 1500
    D = np.tril(rng.integers(0, 40, (n, n)).astype(float), -1)
    D = D + D.T
    merges, heights = upgma(torch.as_tensor(D[None], device=cuda))
    m, h = otb.upgma_vectorised(D)
    assert np.array_equal(merges[0].cpu().numpy(), m)
    assert np.array_equal(heights[0].cpu().numpy(), h)


@pytest.mark.parametrize("kind", ["distinct", "equal"])
def test_sample_trees_match_reference_chain(cuda, kind):
    """alignment -> snp-dists -> sUPGMA correction -> UPGMA -> Newick -> TreeData.from_newick (+ sample times)."""
    from vbsky_b200.treebuild import PackedAlignment, merges_to_newick, sample_trees, snp_dists, supgma_distances, upgma

    rng = np.random.default_rng(5)
    N, L, T, n = 60, 1500, 4, 24
    seqs = _random_alignment(rng, N, L, p_var=0.08)
    names = [f"s{i}" for i in range(N)]
    tm = _times(rng, kind, N)
    sample_times = {names[i]: float(tm[i]) for i in range(N)}
    rows = np.stack([rng.choice(N, n, replace=False) for _ in range(T)])
    aln = PackedAlignment(seqs, cuda)
    tds, maps, omega = sample_trees(aln, names, sample_times, rows)
    r, c = np.tril_indices(n, -1)
    for t in range(T):
        sub = [names[i] for i in rows[t]]
        pw = otb.snp_dists([seqs[i] for i in rows[t]])[r, c]
        mat, om = otb.get_distance_matrix(pw, np.array([sample_times[s] for s in sub]))
        _, _, nested = otb.upgma(mat, sub)
        want, want_map = otd.TreeData.from_newick(otb.nested_to_newick(nested))
        assert omega[t] == pytest.approx(om, rel=1e-9, abs=1e-12)
        got = tds[t]
        assert np.array_equal(got.postorder, want.postorder)
        assert np.array_equal(got.child_parent, want.child_parent)
        assert np.array_equal(got.parent_child, want.parent_child)
        for s in sub:
            assert maps[t][s] == want_map[s]
            assert got.sample_times[maps[t][s]] == sample_times[s]
    # the Newick text round-trips to the same numbering through the product's own reader
    from vbsky_b200.tree_data import TreeData

    dist = snp_dists(aln, rows)
    D, _ = supgma_distances(dist, np.array([[sample_times[names[i]] for i in r_] for r_ in rows]))
    merges, heights = upgma(D)
    nwk = merges_to_newick(merges[0].cpu().numpy(), heights[0].cpu().numpy(), [names[i] for i in rows[0]])
    td2, map2 = TreeData.from_newick(nwk)
    assert np.array_equal(td2.child_parent, tds[0].child_parent)
    assert all(map2[names[i]] == maps[0][names[i]] for i in rows[0])


def test_full_size_ensemble_of_starting_trees(cuda):
    """BASELINE config 3 shape: 100 subsamples of 200 tips from a 29 903-column alignment."""
    from vbsky_b200.treebuild import PackedAlignment, snp_dists, supgma_distances, upgma

    rng = np.random.default_rng(7)
    N, L, T, n = 1000, 29903, 100, 200
    seqs = _random_alignment(rng, N, L, p_var=0.01)
    aln = PackedAlignment(seqs, cuda)
    rows = np.stack([rng.choice(N, n, replace=False) for _ in range(T)])
    times = rng.uniform(0, 1, (T, n))
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    dist = snp_dists(aln, rows)
    e[1].record()
    D, omega = supgma_distances(dist, times)
    e[2].record()
    merges, heights = upgma(D)
    e[3].record()
    torch.cuda.synchronize()
    print("snp_dists %.2f ms, supgma %.2f ms, upgma %.2f ms" % tuple(e[i].elapsed_time(e[i + 1]) for i in range(3)))
    merges, heights, dist = merges.cpu().numpy(), heights.cpu().numpy(), dist.cpu().numpy()
    for t in (0, 57, 99):
        assert np.array_equal(dist[t], otb.snp_dists([seqs[i] for i in rows[t]]))
    for t in range(T):
        assert sorted(merges[t].reshape(-1).tolist()) == list(range(2 * n - 2))
        assert all(max(m) < n + k for k, m in enumerate(merges[t]))
        assert np.all(np.diff(heights[t]) >= -1e-9)
    assert np.all(np.isfinite(omega.cpu().numpy()))
