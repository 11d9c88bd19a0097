"""Seeded synthetic ensembles of the shapes named in BASELINE.json: random
coalescent-shaped trees numbered like ``TreeData.from_nx``, heterochronous sampling times,
sequences evolved down each tree under the substitution model plus a fraction of missing data.
Used by bench.py and the tests; not part of the likelihood path."""
from typing import Any, List, NamedTuple, Optional

import numpy as np

# NumPy only at module level: bench.py's reference arm loads this file by path (without importing the package, so that
# the product's shared library is never mapped into the reference process) and passes the oracle's tree class.
# The package-relative imports below happen lazily, only when the product's classes are wanted.
SubstitutionModel = Any
TreeData = Any


class _Tips(NamedTuple):
    patterns: np.ndarray  # uint8 [S, n] state masks
    counts: np.ndarray    # float64 [S]


def _product_tree_from_edges(edges):
    from .tree_data import TreeData as PTD

    return PTD.from_edges(edges)[0]


def random_tree(n: int, rng: np.random.Generator, caterpillar: bool = False, tree_from_edges=None) -> TreeData:
    """Random binary topology: repeatedly join two uniformly chosen live lineages (Kingman shape),
    or a ladder if ``caterpillar``; sample_times ~ U(0,1) with the most recent tip at 0.
    ``tree_from_edges(edges) -> TreeData``-like (default: the product's ``TreeData.from_edges``)."""
    live = [("t", i) for i in range(n)]
    kids = {}
    k = 0
    while len(live) > 1:
        if caterpillar:
            i, j = 0, 1
        else:
            i, j = rng.choice(len(live), size=2, replace=False)
        a, b = live[i], live[j]
        node = ("i", k)
        k += 1
        kids[node] = (a, b)
        live = [x for idx, x in enumerate(live) if idx not in (i, j)]
        live.insert(0, node) if caterpillar else live.append(node)
    root = live[0]
    edges = []
    stack = [root]
    while stack:  # root-first insertion
        u = stack.pop()
        if u in kids:
            for v in kids[u]:
                edges.append((u, v))
            stack.extend(reversed(kids[u]))
    td = (tree_from_edges or _product_tree_from_edges)(edges)
    st = rng.uniform(0.0, 1.0, n)
    st[rng.integers(n)] = 0.0
    return type(td)(td.postorder, td.child_parent, td.parent_child, st)


class SampledHeights(NamedTuple):
    proportions: np.ndarray      # [K, n-2]
    root_proportion: np.ndarray  # [K]
    heights: np.ndarray          # [K, 2n-1]
    branch_lengths: np.ndarray   # [K, 2n-2]  (time units, not yet times clock rate)


def sample_heights(td: TreeData, K: int, rng: np.random.Generator, origin: float = 0.3) -> SampledHeights:
    """K draws of (proportions, root_proportion) ~ sigmoid(N(0,1)) clipped to [1e-7, 1-1e-7] (the
    range of the reference's ZeroOne flow, transform.py:195) and the implied heights/branch lengths,
    with origin_start = max(sample_times) as fasta.py:372 sets it."""
    n = td.n
    sig = lambda z: np.clip(1.0 / (1.0 + np.exp(-z)), 1e-7, 1 - 1e-7)
    props = sig(rng.standard_normal((K, n - 2)))
    rprop = sig(rng.standard_normal(K))
    hs = np.stack([td.height_transform(rprop[k] * origin, props[k]) for k in range(K)])
    bl = hs[:, td.child_parent[:-1]] - hs[:, :-1]
    return SampledHeights(props, rprop, hs, bl)


def evolve_codes(td: TreeData, bl: np.ndarray, Q: SubstitutionModel, L: int, rng: np.random.Generator,
                 missing: float = 0.01) -> np.ndarray:
    """Evolve L sites from a stationary root down the tree (branch lengths ``bl`` already in
    substitutions/site) -> uint8 [L, n] state masks with a fraction ``missing`` set to N (15)."""
    n, N = td.n, td.N
    state = np.zeros((N, L), dtype=np.uint8)
    state[N - 1] = rng.choice(4, size=L, p=np.asarray(Q.pi) / np.sum(Q.pi))
    for v in range(N - 2, -1, -1):
        P = np.clip(Q.expm(bl[v]), 0.0, 1.0)
        cdf = np.cumsum(P, axis=1)
        cdf[:, -1] = 1.0
        u = rng.random(L, dtype=np.float32)
        ps = state[td.child_parent[v]]
        state[v] = (u[:, None] > cdf[ps][:, :3]).sum(1).astype(np.uint8)
    codes = (np.uint8(1) << state[:n]).T.copy()
    if missing > 0:
        codes[rng.random(codes.shape, dtype=np.float32) < missing] = 15
    return codes


def evolve_codes_cuda(td: TreeData, bl: np.ndarray, Q: SubstitutionModel, L: int, seed: int, device,
                      missing: float = 0.01) -> np.ndarray:
    """Same model as ``evolve_codes`` with the per-node sampling done by torch on the GPU (bench
    set-up only: 100 alignments of 29,903 x 200 take seconds instead of minutes)."""
    import torch

    n, N = td.n, td.N
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    state = torch.zeros((N, L), dtype=torch.int64, device=device)
    pi = torch.as_tensor(np.asarray(Q.pi) / np.sum(Q.pi), device=device)
    state[N - 1] = torch.multinomial(pi, L, replacement=True, generator=gen)
    for v in range(N - 2, -1, -1):
        cdf = torch.as_tensor(np.cumsum(np.clip(Q.expm(bl[v]), 0.0, 1.0), axis=1), device=device)
        u = torch.rand(L, device=device, dtype=torch.float64, generator=gen)
        ps = state[int(td.child_parent[v])]
        state[v] = (u[:, None] > cdf[ps][:, :3]).sum(1)
    codes = (torch.ones((), dtype=torch.int64, device=device) << state[:n]).T.contiguous()
    if missing > 0:
        codes[torch.rand(codes.shape, device=device, generator=gen) < missing] = 15
    return codes.to(torch.uint8).cpu().numpy()


class SyntheticEnsemble(NamedTuple):
    tds: List[TreeData]
    tips: List[Any]  # PackedTips-like (patterns, counts)
    heights: List[SampledHeights]
    clock_rate: float
    origin: float


def make_ensemble(T: int, n: int, S: int, K: int, seed: int, Q: SubstitutionModel, clock_rate: float = 1.12e-3,
                  origin: float = 0.3, compress: bool = False, missing: float = 0.01,
                  device=None, tree_from_edges=None) -> SyntheticEnsemble:
    """T trees x n tips x S sites (used directly as patterns with count 1 unless ``compress``),
    K height samples per tree.  ``device``: evolve the sequences with torch on that CUDA device (a different random
    stream from the NumPy one; the trees and heights are the same either way)."""
    rng = np.random.default_rng(seed)
    tds, tips, hts = [], [], []
    if tree_from_edges is None:
        from .tips import PackedTips
    else:
        PackedTips = _Tips
    # trees and heights first (one stream), sequences from a per-tree stream: the first t trees, their heights and -- on
    # the NumPy path -- their sequences do not depend on T, so a subset of the ensemble can be regenerated on its own
    for t in range(T):
        td = random_tree(n, rng, tree_from_edges=tree_from_edges)
        tds.append(td), hts.append(sample_heights(td, K, rng, origin))
    for t in range(T):
        td, sh = tds[t], hts[t]
        if device is not None:
            codes = evolve_codes_cuda(td, sh.branch_lengths[0] * clock_rate, Q, S, seed * 100003 + t, device, missing)
        else:
            codes = evolve_codes(td, sh.branch_lengths[0] * clock_rate, Q, S, np.random.default_rng([seed, t]), missing)
        if compress:
            from .tips import compress_patterns

            pat, cnt = compress_patterns(codes)
            tp = PackedTips(pat, cnt.astype(np.float64))
        else:
            tp = PackedTips(codes, np.ones(S))
        tips.append(tp)
    if compress:
        from .tips import pad_tips

        tips = pad_tips(tips)
    return SyntheticEnsemble(tds, tips, hts, clock_rate, origin)
