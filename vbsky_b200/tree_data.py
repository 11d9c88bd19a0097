"""Host-side tree arrays: the ``TreeData`` container the reference passes into its kernels
(/root/reference/vbsky/tree_data.py:17-23) and the constructors that define its node numbering.

Same names, argument meaning and results as the reference (``from_nx`` numbering:
tree_data.py:464-507; derived arrays :55-117; height transforms :256-376), implemented with
plain array passes instead of JAX scans so that nothing here needs JAX, networkx or Bio.
The device-side consumers are ``vbsky_layout_create`` (schedules) and the heights kernels.
"""
from typing import Dict, Hashable, Iterable, List, NamedTuple, Optional, Sequence, Tuple

import numpy as np


class _OrderedDiGraph:
    """Minimal insertion-ordered digraph (the subset of the networkx API ``from_nx`` reads)."""

    def __init__(self):
        self.succ: Dict[Hashable, Dict[Hashable, dict]] = {}
        self.pred: Dict[Hashable, Dict[Hashable, dict]] = {}

    def add_node(self, u):
        if u not in self.succ:
            self.succ[u] = {}
            self.pred[u] = {}

    def add_edge(self, u, v, **attr):
        self.add_node(u)
        self.add_node(v)
        d = self.succ[u].setdefault(v, {})
        d.update(attr)
        self.pred[v][u] = d

    def remove_edge(self, u, v):
        del self.succ[u][v]
        del self.pred[v][u]

    @property
    def nodes(self):
        return list(self.succ)

    def __getitem__(self, u):
        return self.succ[u]


def _postorder(nodes: Sequence, succ) -> List:
    """Depth-first postorder over all nodes in container order, following successors in insertion
    order -- the visiting order of ``networkx.dfs_postorder_nodes(G)`` with no source."""
    seen = set()
    out = []
    for start in nodes:
        if start in seen:
            continue
        seen.add(start)
        stack = [(start, iter(succ[start]))]
        while stack:
            node, it = stack[-1]
            for child in it:
                if child not in seen:
                    seen.add(child)
                    stack.append((child, iter(succ[child])))
                    break
            else:
                stack.pop()
                out.append(node)
    return out


class TreeData(NamedTuple):
    """Topology of a rooted binary tree with n tips (tree_data.py:17-23).

    postorder    [2n-1]    a postorder traversal (root last)
    child_parent [2n-1]    parent of each node, -1 for the root
    parent_child [2n-1,2]  children of each node ([larger index, smaller index]); -1,-1 for tips
    sample_times [n]       sampling time of each tip (0 = most recent)
    """

    postorder: np.ndarray
    child_parent: np.ndarray
    parent_child: np.ndarray
    sample_times: np.ndarray

    # -- derived quantities (tree_data.py:44-117) ------------------------------------------------
    @property
    def N(self) -> int:
        return len(self.postorder)

    @property
    def n(self) -> int:
        return (self.N + 1) // 2

    @property
    def leaves(self):
        return np.arange(self.n)

    @property
    def internal_nodes(self):
        return np.arange(self.n, self.N)

    @property
    def preorder(self):
        return self.postorder[::-1]

    def root(self):
        return self.preorder[0]

    @property
    def siblings(self) -> np.ndarray:
        sib = np.full(self.N, -1, dtype=np.int64)
        kids = self.parent_child[self.n :]
        sib[kids[:, 0]] = kids[:, 1]
        sib[kids[:, 1]] = kids[:, 0]
        return sib

    @property
    def lower_sampling_times(self) -> np.ndarray:
        """Most recent sampling time below each node (tree_data.py:80-117)."""
        low = np.zeros(self.N, dtype=np.float64)
        low[: self.n] = self.sample_times
        for v in range(self.n, self.N):  # children always precede parents in the numbering
            a, b = self.parent_child[v]
            low[v] = max(low[a], low[b])
        return low

    # -- height parametrisation (tree_data.py:256-367) --------------------------------------------
    def height_transform(self, root_height, proportions) -> np.ndarray:
        """heights[i] = (1-p_i) * lower_sampling_times[i] + p_i * heights[parent(i)] for internal
        non-root i; tips sit at their sampling times; root at root_height + max(sample_times)."""
        proportions = np.asarray(proportions, dtype=np.float64)
        if len(proportions) != self.N - self.n - 1:
            raise AssertionError("one proportion per internal non-root node is required")
        lst = self.lower_sampling_times
        h = np.zeros(self.N, dtype=np.float64)
        h[: self.n] = lst[: self.n]
        h[self.N - 1] = float(np.asarray(root_height).reshape(-1)[0]) + self.sample_times.max()
        for v in range(self.N - 2, self.n - 1, -1):  # parents (larger index) before children
            p = proportions[v - self.n]
            h[v] = (1.0 - p) * lst[v] + p * h[self.child_parent[v]]
        return h

    def inverse_height_transform(self, heights) -> Tuple[float, np.ndarray]:
        heights = np.asarray(heights, dtype=np.float64)
        if len(heights) != self.N:
            raise AssertionError("one height per node is required")
        lst = self.lower_sampling_times
        inner = np.arange(self.n, self.N - 1)
        hp = heights[self.child_parent[inner]]
        return heights[self.N - 1], (heights[inner] - lst[inner]) / (hp - lst[inner])

    def bl_to_nh(self, blens) -> np.ndarray:
        """Node heights from the branch length above each node (tree_data.py:369-376)."""
        h = np.zeros(self.N)
        h[: self.n] = self.sample_times
        for i in self.postorder[:-1]:
            h[self.child_parent[i]] = h[i] + blens[i]
        return h

    def branch_lengths(self, heights) -> np.ndarray:
        """heights[parent] - heights[node] for every non-root node (bdsky.py:307)."""
        heights = np.asarray(heights)
        return heights[self.child_parent[:-1]] - heights[:-1]

    def to_newick(self, branch_lengths=None, node_mapping: Optional[dict] = None) -> str:
        node_mapping = node_mapping or {}

        def label(s, i):
            if branch_lengths is None or i >= len(branch_lengths):
                return s
            return f"{s}:{branch_lengths[i]}"

        out: Dict[int, str] = {}
        for v in range(self.N):  # children precede parents
            if v < self.n:
                out[v] = label(str(node_mapping.get(v, v)), v)
            else:
                a, b = self.parent_child[v]
                out[v] = label(f"({out.pop(int(a))},{out.pop(int(b))})", v)
        return out[self.N - 1]

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def from_nx(cls, G, branch_length_key: str = "branch_length") -> Tuple["TreeData", dict]:
        """Number a rooted binary tree given as a directed graph (any object with networkx's
        ``nodes`` / ``G[u]`` / ``pred`` interface).  Tips get 0..n-1 and internal nodes n..2n-2, each
        in depth-first postorder encounter order, which makes every child index smaller than its
        parent's and the root 2n-2.  Edge attribute ``branch_length_key`` (default 1 when missing)
        gives branch lengths; the tip furthest from the root is sampled at time 0."""
        nodes = list(G.nodes)
        succ = {u: list(G[u]) for u in nodes}
        pred = {u: list(G.pred[u]) for u in nodes}
        is_tip = {u: len(succ[u]) == 0 and len(pred[u]) == 1 for u in nodes}
        n = sum(is_tip.values())
        order = _postorder(nodes, succ)
        remap = {}
        next_tip, next_inner = 0, n
        for u in order:
            if is_tip[u]:
                remap[u] = next_tip
                next_tip += 1
            else:
                remap[u] = next_inner
                next_inner += 1
        M = len(nodes)
        if M != 2 * n - 1 or next_inner != M:
            raise ValueError("graph is not a rooted binary tree")
        postorder = np.array([remap[u] for u in order])

        roots = [u for u in nodes if len(succ[u]) == 2 and len(pred[u]) == 0]
        if len(roots) != 1:
            raise ValueError("graph must have exactly one root of out-degree 2")
        # distance from the root, accumulated parent -> child like Dijkstra does on a tree
        dist = {roots[0]: 0}
        stack = [roots[0]]
        while stack:
            u = stack.pop()
            for v in succ[u]:
                dist[v] = dist[u] + G[u][v].get(branch_length_key, 1)
                stack.append(v)
        far = max(dist[u] for u in nodes if is_tip[u])
        inv = {i: u for u, i in remap.items()}
        sample_times = np.array([far - dist[inv[i]] for i in range(n)])

        child_parent = np.full(M, -1)
        for u in nodes:
            if pred[u]:
                child_parent[remap[u]] = remap[pred[u][0]]
        parent_child = np.full((M, 2), -1)
        for i in range(M):  # ascending child index: column 0 ends up holding the larger child
            pa = child_parent[i]
            if pa >= 0:
                parent_child[pa, 1] = parent_child[pa, 0]
                parent_child[pa, 0] = i
        return cls(postorder, child_parent, parent_child, sample_times), remap

    @classmethod
    def from_edges(cls, edges: Iterable[Tuple[Hashable, Hashable]], lengths: Optional[Iterable[float]] = None):
        """Convenience: build from (parent, child) pairs in insertion order (+ optional lengths)."""
        G = _OrderedDiGraph()
        if lengths is None:
            for u, v in edges:
                G.add_edge(u, v)
        else:
            for (u, v), w in zip(edges, lengths):
                G.add_edge(u, v, branch_length=w)
        return cls.from_nx(G)

    @classmethod
    def from_newick(cls, newick: str, return_branch_lengths: bool = False):
        """Newick string -> (TreeData, name->index mapping[, branch lengths]) with the reference's
        conventions (tree_data.py:149-254): only tips are named, a missing (or zero) branch length
        counts as 1, multifurcations are resolved into a ladder with unit-length new edges."""
        G, names = _parse_newick(newick)
        # resolve multifurcations exactly as the reference's loop does (first child stays, the rest
        # are chained below freshly created nodes; the new edges carry no length -> default 1)
        fresh = 0
        for u in [u for u in G.nodes if len(G[u]) > 2]:
            kids = list(G[u])
            anchor = u
            for v in kids[1:-1]:
                nn = ("poly", fresh)
                fresh += 1
                names[nn] = f"poly{fresh}"
                G.add_edge(anchor, nn)
                G.add_edge(nn, v)
                G.remove_edge(u, v)
                anchor = nn
            v = kids[-1]
            G.add_edge(anchor, v)
            G.remove_edge(u, v)
        td, remap = cls.from_nx(G, "weight")
        mapping = {names[u]: i for u, i in remap.items()}
        if return_branch_lengths:
            bl = np.zeros(2 * td.n - 2)
            for u, i in remap.items():
                if G.pred[u]:
                    pa = next(iter(G.pred[u]))
                    bl[i] = G[pa][u].get("weight", 1)
            return td, mapping, bl
        return td, mapping


def _parse_newick(s: str):
    """Newick -> (_OrderedDiGraph in parent-first / file order with 'weight' edge attributes,
    node -> name).  Unnamed clades are called internal0, internal1, ... in parent-first order and a
    missing or zero branch length counts as 1 (tree_data.py:198-206)."""
    s = s.strip().rstrip(";").strip()
    if not s.startswith("("):
        raise ValueError("newick must start with '('")
    G = _OrderedDiGraph()
    names, lengths = {}, {}
    pos = 1
    count = 1
    root = ("node", 0)
    G.add_node(root)
    open_nodes = [root]

    def read_label():
        nonlocal pos
        st = pos
        while pos < len(s) and s[pos] not in ",():":
            pos += 1
        name = s[st:pos].strip() or None
        bl = None
        if pos < len(s) and s[pos] == ":":
            pos += 1
            st = pos
            while pos < len(s) and s[pos] not in ",()":
                pos += 1
            bl = float(s[st:pos])
        return name, bl

    expect_child = True
    while open_nodes:
        if expect_child:
            child = ("node", count)
            count += 1
            G.add_edge(open_nodes[-1], child)
            if pos < len(s) and s[pos] == "(":
                pos += 1
                open_nodes.append(child)
                continue
            names[child], lengths[child] = read_label()
            expect_child = False
        else:
            ch = s[pos] if pos < len(s) else None
            if ch == ",":
                pos += 1
                expect_child = True
            elif ch == ")":
                pos += 1
                done = open_nodes.pop()
                names[done], lengths[done] = read_label()
            else:
                raise ValueError(f"malformed newick at position {pos}")
    anon = 0
    for u in G.nodes:
        if names.get(u) is None:
            names[u] = f"internal{anon}"
            anon += 1
        for p in G.pred[u]:
            G.succ[p][u]["weight"] = lengths.get(u) or 1
    return G, names
