"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's starting-tree construction
(`SeqData.sample_trees`, /root/reference/vbsky/fasta.py:203-296): pairwise SNP distances, the serial-sample
UPGMA distance correction and the UPGMA clustering.  Never imported by the product path.

Two third-party pieces of this path are absent from /root/reference and from this image; their published
algorithms are restated here from their documentation and the reference's call sites (PARITY UNPINNED for
both -- no golden vectors exist in the reference):

* ``snp-dists`` (tseemann/snp-dists, unpinned by the reference; called as ``snp-dists -b temp.fa``,
  fasta.py:268-274).  Default options: sequences are upper-cased, every character outside ``AGTC`` is
  replaced by an ignore character, and the distance of two sequences is the number of columns where the
  characters differ and neither is ignored.
* ``Bio.Phylo.TreeConstruction.DistanceTreeConstructor.upgma`` (biopython>=1.79, setup.py:9; called at
  fasta.py:283).  Lower-triangular list-of-lists matrix; every round scans ``for i in 1..len-1: for j in
  0..i-1`` keeping the LAST entry with ``min_dist >= dm[i, j]``; the two clades become children (clade
  min_i first, min_j second) of a new ``Inner<k>`` clade stored at position min_j; the new row is the plain
  average ``(dm[min_i, k] + dm[min_j, k]) * 1.0 / 2`` (no cluster-size weights); row/column min_i is deleted.

``get_distance_matrix`` follows /root/reference/vbsky/upgma.py:25-62 line by line (sklearn's
``LinearRegression`` is available here and is used as the reference uses it).
"""
from typing import List, Sequence, Tuple

import numpy as np


def snp_dists(seqs: Sequence[str]) -> np.ndarray:
    """[n, n] int64 matrix ``snp-dists -b`` prints (fasta.py:268-281 reads its lower triangle)."""
    a = np.array([np.frombuffer(s.upper().encode("ascii"), dtype=np.uint8) for s in seqs])
    ok = np.isin(a, np.frombuffer(b"AGTC", dtype=np.uint8))
    n = len(seqs)
    out = np.zeros((n, n), dtype=np.int64)
    for i in range(n):
        d = (a[i][None, :] != a) & ok[i][None, :] & ok
        out[i] = d.sum(axis=1)
    return out


def get_distance_matrix(pw_dist: np.ndarray, times: np.ndarray, single_theta: bool = False) -> Tuple[np.ndarray, float]:
    """upgma.py:25-62: pw_dist = lower triangle in np.tril_indices(n, -1) order.  Returns the [n, n] array
    ``mat`` whose strict lower triangle the reference hands to DistanceMatrix, and omega."""
    from sklearn.linear_model import LinearRegression

    times = np.asarray(times, dtype=float)
    n = len(times)
    nc2 = int(n * (n - 1) / 2)
    r, c = np.tril_indices(n, -1)
    y = pw_dist
    dt = np.abs(times[:, None] - times)[(r, c)]
    uniq_times = np.sort(np.unique(times))[::-1]
    times_dict = {uniq_times[i]: i for i in range(len(uniq_times))}
    if single_theta:
        reg = LinearRegression(fit_intercept=True).fit(dt.reshape(-1, 1), y)
    else:
        larger = np.where(times[r] >= times[c], r, c)
        lsm = np.array([times_dict[times[ell]] for ell in larger])
        X = np.zeros((nc2, len(times_dict)))
        X[(np.arange(nc2), lsm)] = 1
        X = np.column_stack((X, dt))
        reg = LinearRegression(fit_intercept=False).fit(X, y)
    omega = reg.coef_[-1]
    adjusted_y = y + omega * (times[r] + times[c] - 2 * uniq_times[-1])
    mat = np.zeros((n, n))
    mat[(r, c)] = adjusted_y
    return mat, float(omega)


class _LowerTri:
    """The slice of Bio.Phylo.TreeConstruction.DistanceMatrix that ``upgma`` touches."""

    def __init__(self, names: List, mat: np.ndarray):
        self.names = list(names)
        self.matrix = [[float(mat[i, j]) for j in range(i)] + [0.0] for i in range(len(names))]

    def __len__(self):
        return len(self.names)

    def get(self, i, j):
        return self.matrix[i][j] if i > j else self.matrix[j][i]

    def set(self, i, j, v):
        if i > j:
            self.matrix[i][j] = v
        else:
            self.matrix[j][i] = v

    def delete(self, index):
        for i in range(index + 1, len(self)):
            del self.matrix[i][index]
        del self.matrix[index]
        del self.names[index]


def upgma(mat: np.ndarray, names: Sequence = None):
    """Restated DistanceTreeConstructor.upgma.  ``mat``: [n, n], strict lower triangle used.

    Returns (merges, heights, tree): merges[k] = (clade1, clade2) of the k-th inner clade, where a clade is a
    tip index < n or n + (index of an earlier merge); heights[k] = min_dist / 2; tree = nested tuples of tip
    names in child order (clade1, clade2)."""
    n = mat.shape[0]
    names = list(range(n)) if names is None else list(names)
    dm = _LowerTri(list(range(n)), mat)
    clades = [(i, names[i]) for i in range(n)]  # (clade id, nested tuple)
    merges, heights = [], []
    min_i = min_j = 0
    while len(dm) > 1:
        min_dist = dm.get(1, 0)
        for i in range(1, len(dm)):
            for j in range(0, i):
                if min_dist >= dm.get(i, j):
                    min_dist = dm.get(i, j)
                    min_i, min_j = i, j
        c1, c2 = clades[min_i], clades[min_j]
        inner = (n + len(merges), (c1[1], c2[1]))
        merges.append((c1[0], c2[0]))
        heights.append(min_dist * 1.0 / 2)
        clades[min_j] = inner
        del clades[min_i]
        for k in range(0, len(dm)):
            if k != min_i and k != min_j:
                dm.set(min_j, k, (dm.get(min_i, k) + dm.get(min_j, k)) * 1.0 / 2)
        dm.delete(min_i)
    return np.array(merges, dtype=np.int32).reshape(-1, 2), np.array(heights), clades[0][1]


def nested_to_newick(tree) -> str:
    """Nested tuples of tip names -> Newick without lengths (what TreeData.from_newick needs: the reference
    discards the UPGMA branch lengths, fasta.py:529-537 overwrites sample_times)."""
    def f(t):
        if isinstance(t, tuple):
            return "(" + ",".join(f(c) for c in t) + ")"
        return str(t)
    return f(tree) + ";"


def upgma_vectorised(mat: np.ndarray):
    """Same merges and heights as ``upgma`` (checked against it in tests/test_treebuild_oracle.py) with NumPy doing each
    round's scan: live clades keep their relative order, so Biopython's positions are the live slots in ascending order
    and its last ``>=`` hit is the last minimum in row-major order over the live lower triangle.  For n in the thousands,
    where the literal list-of-lists loop is too slow to run in a test."""
    n = mat.shape[0]
    D = np.tril(np.asarray(mat, dtype=np.float64), -1)
    D = D + D.T
    alive = np.arange(n)
    cid = np.arange(n)
    merges, heights = [], []
    for rnd in range(n - 1):
        sub = D[np.ix_(alive, alive)]
        m = len(alive)
        r, c = np.tril_indices(m, -1)
        vals = sub[r, c]
        k = np.flatnonzero(vals == vals.min())[-1]
        pi, pj = r[k], c[k]
        i, j = alive[pi], alive[pj]
        d = vals[k]
        merges.append((cid[i], cid[j]))
        heights.append(d * 1.0 / 2)
        others = alive[(alive != i) & (alive != j)]
        new = (D[i, others] + D[j, others]) * 1.0 / 2
        D[j, others] = new
        D[others, j] = new
        cid[j] = n + rnd
        alive = np.delete(alive, pi)
    return np.array(merges, dtype=np.int32).reshape(-1, 2), np.array(heights)
