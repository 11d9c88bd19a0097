"""Starting trees for an ensemble, built on the device: the host side of ``SeqData.sample_trees``
(/root/reference/vbsky/fasta.py:203-296) for the default (non-audacity) branch.

The reference runs, per subsample, the external ``snp-dists`` program, ``get_distance_matrix``
(upgma.py:25-62, an sklearn regression) and Biopython's ``DistanceTreeConstructor.upgma`` (pure Python,
O(n^3)), writes Newick and re-reads it with ``TreeData.from_newick`` (fasta.py:305-311, 529-537).  Here the
alignment is packed into bit planes once and the three stages run as batched kernels over all T subsamples
(``csrc/treebuild.cu``); only the O(n) renumbering into ``TreeData`` stays on the host.
"""
import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .tree_data import TreeData, _OrderedDiGraph


def _stream(dev):
    import torch

    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


class PackedAlignment:
    """An alignment resident on the device: raw characters [N, L] and the A/C/G/T bit planes."""

    def __init__(self, seqs: Sequence[str], device):
        import torch

        seqs = [s if isinstance(s, str) else "".join(s) for s in seqs]
        self.N, self.L = len(seqs), len(seqs[0])
        if any(len(s) != self.L for s in seqs):
            raise ValueError("sequences must have equal length")
        self.device = torch.device(device)
        self.chars = torch.frombuffer(bytearray("".join(seqs).encode("latin-1")), dtype=torch.uint8).to(self.device).view(self.N, self.L)
        nbytes = _lib.lib.vbsky_snp_planes_bytes(self.N, self.L)
        self.planes = torch.empty(nbytes // 4, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.vbsky_snp_pack_device(_p(self.chars), self.N, self.L, _p(self.planes), _stream(self.device)))


def snp_dists(aln: PackedAlignment, rows):
    """``snp-dists -b`` of every subsample: ``rows`` int [T, n] indices into the alignment -> CUDA int32 [T, n, n]."""
    import torch

    rows = torch.as_tensor(rows, dtype=torch.int32, device=aln.device).contiguous()
    if rows.dim() == 1:
        rows = rows[None]
    if rows.numel() and (int(rows.min()) < 0 or int(rows.max()) >= aln.N):
        raise IndexError("subsample row outside the alignment")
    T, n = rows.shape
    dist = torch.empty((T, n, n), dtype=torch.int32, device=aln.device)
    with torch.cuda.device(aln.device):
        _lib.check(_lib.lib.vbsky_snp_dists_device(_p(aln.planes), aln.N, aln.L, _p(rows), T, n, _p(dist), _stream(aln.device)))
    return dist


def supgma_distances(dist, times, single_theta: bool = False):
    """``get_distance_matrix`` (upgma.py:25-62) batched: dist CUDA int32 [T, n, n], times [T, n] ->
    (D CUDA float64 [T, n, n] symmetric, omega CUDA float64 [T])."""
    import torch

    dist = dist.contiguous()
    T, n, _ = dist.shape
    times = torch.as_tensor(times, dtype=torch.float64, device=dist.device).contiguous().view(T, n)
    D = torch.empty((T, n, n), dtype=torch.float64, device=dist.device)
    omega = torch.empty(T, dtype=torch.float64, device=dist.device)
    with torch.cuda.device(dist.device):
        _lib.check(_lib.lib.vbsky_supgma_distances_device(T, n, _p(dist), _p(times), int(bool(single_theta)), _p(D), _p(omega), _stream(dist.device)))
    return D, omega


def upgma(D):
    """``DistanceTreeConstructor.upgma`` batched: D CUDA float64 [T, n, n] (lower triangle is what Biopython
    reads; must be symmetric) -> (merges CUDA int32 [T, n-1, 2], heights CUDA float64 [T, n-1])."""
    import torch

    T, n, _ = D.shape
    work = D.clone().contiguous()
    merges = torch.empty((T, n - 1, 2), dtype=torch.int32, device=D.device)
    heights = torch.empty((T, n - 1), dtype=torch.float64, device=D.device)
    with torch.cuda.device(D.device):
        _lib.check(_lib.lib.vbsky_upgma_device(T, n, _p(work), _p(merges), _p(heights), _stream(D.device)))
    return merges, heights


def _children(merges: np.ndarray):
    n = merges.shape[0] + 1
    return n, {n + k: (int(merges[k, 0]), int(merges[k, 1])) for k in range(n - 1)}


def merges_to_tree(merges: np.ndarray, names: Sequence, sample_times: Optional[Dict] = None) -> Tuple[TreeData, Dict]:
    """One subsample's merge list -> (TreeData, node_mapping) exactly as writing Biopython's tree to Newick and
    reading it back with ``TreeData.from_newick`` numbers it (children in (clade1, clade2) order, inner clades
    named ``Inner<k>``); with ``sample_times`` (name -> time) the tips' times are filled in as
    ``_process_tree`` does (fasta.py:529-537)."""
    merges = np.asarray(merges)
    n, kids = _children(merges)
    label = lambda c: names[c] if c < n else f"Inner{c - n + 1}"
    G = _OrderedDiGraph()
    stack = [2 * n - 2]
    G.add_node(label(2 * n - 2))
    while stack:  # parent-first, first child's subtree before the second child (Phylo.to_networkx order)
        u = stack.pop()
        if u < n:
            continue
        a, b = kids[u]
        G.add_edge(label(u), label(a))
        G.add_edge(label(u), label(b))
        stack.append(b)
        stack.append(a)
    td, remap = TreeData.from_nx(G)
    td = td._replace(sample_times=np.asarray(td.sample_times, dtype=np.float64))
    if sample_times is not None:
        inv = {i: u for u, i in remap.items()}
        td.sample_times[:] = [sample_times[inv[i]] for i in range(td.n)]
    return td, dict(remap)


def merges_to_newick(merges: np.ndarray, heights: np.ndarray, names: Sequence) -> str:
    """Newick of one UPGMA tree with ultrametric branch lengths (parent height - child height, %1.5f like
    ``Phylo.write``).  The reference only keeps the topology of this file (fasta.py:529-537)."""
    merges = np.asarray(merges)
    n, kids = _children(merges)
    h = lambda c: 0.0 if c < n else float(heights[c - n])

    def f(c, parent_h):
        bl = "" if parent_h is None else ":%1.5f" % (parent_h - h(c))
        if c < n:
            return f"{names[c]}{bl}"
        a, b = kids[c]
        return f"({f(a, h(c))},{f(b, h(c))})Inner{c - n + 1}{bl}"

    import sys

    sys.setrecursionlimit(max(sys.getrecursionlimit(), 4 * n + 100))
    return f(2 * n - 2, None) + ";"


def sample_trees(aln: PackedAlignment, names: Sequence[str], sample_times: Dict[str, float], rows, single_theta: bool = False):
    """Batched ``SeqData.sample_trees`` + ``process_trees`` for given subsamples: ``rows`` [T, n] indices into
    the alignment (the reference draws them with ``np.random.choice``, fasta.py:253).  Returns
    (tds, node_mappings, omega [T] ndarray)."""
    import torch

    rows_h = np.asarray(torch.as_tensor(rows).cpu(), dtype=np.int64)
    if rows_h.ndim == 1:
        rows_h = rows_h[None]
    times = np.array([[sample_times[names[i]] for i in r] for r in rows_h], dtype=np.float64)
    dist = snp_dists(aln, rows_h)
    D, omega = supgma_distances(dist, times, single_theta)
    merges, _ = upgma(D)
    merges_h = merges.cpu().numpy()
    tds: List[TreeData] = []
    maps: List[Dict] = []
    for t, r in enumerate(rows_h):
        td, nm = merges_to_tree(merges_h[t], [names[i] for i in r], sample_times)
        tds.append(td)
        maps.append(nm)
    return tds, maps, omega.cpu().numpy()
