"""Oracle restatement of /root/reference/vbsky/tree_data.py (test infrastructure).

Follows tree_data.py:17-117 (TreeData and derived arrays), :256-376 (height transforms),
:411-517 (from_nx numbering), :149-254 (from_newick; Bio.Phylo is absent here, so a small
Newick reader rebuilds the same DiGraph that ``Bio.Phylo.to_networkx`` produces -- root first,
then every child in file order, depth first -- and the reference's relabelling / polytomy
breaking is applied to it unchanged in order).  ``lax.scan`` loops become Python loops.
"""
import itertools
from typing import NamedTuple

import networkx as nx
import numpy as np


class TreeData(NamedTuple):
    """tree_data.py:17-23."""

    postorder: np.ndarray
    child_parent: np.ndarray
    parent_child: np.ndarray
    sample_times: np.ndarray

    @property
    def N(self):
        return len(self.postorder)  # tree_data.py:61-62

    @property
    def n(self):
        return (self.N + 1) // 2  # tree_data.py:56-58

    @property
    def preorder(self):
        return self.postorder[::-1]  # tree_data.py:65-66

    @property
    def siblings(self):
        # tree_data.py:69-78.  Leaves' (-1,-1) rows write siblings[-1] = -1, as in the reference.
        siblings = np.zeros(self.N, dtype=int)
        for i in 0, 1:
            siblings[self.parent_child[:, i]] = self.parent_child[:, 1 - i]
        return siblings

    @property
    def lower_sampling_times(self):
        # tree_data.py:81-117
        lowers = np.concatenate([self.sample_times, np.zeros(self.n - 1)]).astype(np.float64)
        for i in self.postorder:
            if i >= self.n:
                lowers[i] = lowers[self.parent_child[i]].max()
        return lowers

    def to_newick(self, branch_lengths=None, node_mapping={}):
        # tree_data.py:119-147
        def _bl(s, i):
            try:
                return s + ":" + str(branch_lengths[i])
            except Exception:
                return s

        def _f(i):
            if i < self.n:
                return _bl(str(node_mapping.get(i, i)), i)
            s = ",".join(map(_f, self.parent_child[i]))
            return _bl("(" + s + ")", i)

        return _f(self.preorder[0])

    def inverse_height_transform(self, heights):
        # tree_data.py:256-285
        heights = np.asarray(heights, dtype=np.float64)
        assert len(heights) == self.N
        root_height = heights[self.preorder[0]]
        internal_nodes = self.preorder[self.preorder >= self.n][1:]
        h = heights[internal_nodes]
        h_d = self.lower_sampling_times[internal_nodes]
        h_p = heights[self.child_parent[internal_nodes]]
        proportions = (h - h_d) / (h_p - h_d)
        proportions = proportions[internal_nodes.argsort()]
        return root_height, proportions

    def height_transform(self, root_height, proportions):
        # tree_data.py:287-367
        proportions = np.asarray(proportions, dtype=np.float64)
        assert len(proportions) == self.N - self.n - 1
        lst = self.lower_sampling_times
        heights = np.concatenate(
            [lst[: self.n], np.zeros(self.n - 2), np.atleast_1d(root_height) + self.sample_times.max()]
        ).astype(np.float64)
        if self.N - self.n > 1:
            for i in self.preorder[1:]:
                if i >= self.n:
                    p = proportions[i - self.n]
                    heights[i] = (1.0 - p) * lst[i] + p * heights[self.child_parent[i]]
        return heights

    def bl_to_nh(self, blens):
        # tree_data.py:369-376
        ret = np.zeros(2 * self.n - 1)
        ret[: self.n] = self.sample_times
        for i in self.postorder[:-1]:
            ret[self.child_parent[i]] = ret[i] + blens[i]
        return ret

    @classmethod
    def from_nx(cls, G: nx.DiGraph, branch_length_key: str = "branch_length"):
        # tree_data.py:411-517
        def is_leaf(x):
            return G.out_degree(x) == 0 and G.in_degree(x) == 1

        n = sum(map(is_leaf, G.nodes()))
        leaf_iter = iter(range(n))
        internal_iter = iter(range(n, 2 * n - 1))
        node_remap = {}
        for x in nx.dfs_postorder_nodes(G):
            node_remap[x] = next(leaf_iter) if is_leaf(x) else next(internal_iter)

        G = nx.relabel_nodes(G, mapping=node_remap)
        postorder = np.array(list(nx.dfs_postorder_nodes(G)))
        M = len(G)

        root = next(x for x in G.nodes if G.out_degree(x) == 2 and G.in_degree(x) == 0)
        leaves = list(filter(is_leaf, G.nodes))
        root_distance = nx.shortest_path_length(G, source=root, weight=branch_length_key)
        m = max(root_distance[k] for k in leaves)
        sample_times = np.array([m - root_distance[x] for x in range(n)])

        child_parent = np.full(M, -1)
        for k in G.pred:
            try:
                child_parent[k] = next(iter(G.pred[k]))
            except StopIteration:
                pass
        parent_child = np.zeros([M, 2], int) - 1
        for i in range(M):
            parent = child_parent[i]
            if parent == -1:
                continue
            parent_child[parent, 1] = parent_child[parent, 0]
            parent_child[parent, 0] = i
        return cls(postorder, child_parent, parent_child, sample_times), node_remap

    @classmethod
    def from_newick(cls, newick: str, return_branch_lengths: bool = False):
        # tree_data.py:149-254 (outgroup pruning omitted: it is Bio.Phylo functionality).
        G, info = _newick_to_digraph(newick)
        i = itertools.count()
        mapping = {}
        for clade in G.nodes:
            name, bl = info[clade]
            mapping[clade] = (name if name is not None else f"internal{next(i)}", bl or 1)
        mapping_names = {k: v[0] for k, v in mapping.items()}
        mapping_blens = dict(mapping.values())
        G1 = nx.relabel_nodes(G, mapping_names)
        for k in mapping_blens:
            try:
                u, v = next(iter(G1.in_edges(k)))
                G1[u][v]["weight"] = mapping_blens[k]
            except StopIteration:
                pass
        poly = [u for u in G1.nodes if G1.out_degree(u) > 2]
        for u in poly:
            edges = list(G1.out_edges(u))
            n0 = u
            for _, v in edges[1:-1]:
                nn = f"poly{next(i)}"
                G1.add_edge(n0, nn)
                G1.add_edge(nn, v)
                G1.remove_edge(u, v)
                n0 = nn
            v = edges[-1][1]
            G1.add_edge(n0, v)
            G1.remove_edge(u, v)
        ret = td, nm = cls.from_nx(G1, "weight")
        if return_branch_lengths:
            q = [(nd, nm[nd]) for nd in G1.nodes]
            bl = np.zeros(2 * td.n - 2)
            while q:
                nd, j = q.pop()
                try:
                    (u, v) = next(iter(G1.in_edges(nd)))
                    bl[j] = G1[u][v]["weight"]
                    q.append((u, td.child_parent[j]))
                except StopIteration:
                    pass
            ret = ret + (bl,)
        return ret


def _newick_to_digraph(newick: str):
    """Minimal Newick reader -> DiGraph with the node/edge insertion order of Bio.Phylo.to_networkx
    (add root; for each child in order: add node, add edge parent->child, recurse)."""
    s = newick.strip().rstrip(";").strip()
    pos = 0
    counter = itertools.count()
    info = {}
    children = {}

    def parse():
        nonlocal pos
        node = next(counter)
        kids = []
        if pos < len(s) and s[pos] == "(":
            pos += 1
            while True:
                kids.append(parse())
                if s[pos] == ",":
                    pos += 1
                    continue
                assert s[pos] == ")", f"bad newick at {pos}"
                pos += 1
                break
        start = pos
        while pos < len(s) and s[pos] not in ",():":
            pos += 1
        name = s[start:pos].strip() or None
        bl = None
        if pos < len(s) and s[pos] == ":":
            pos += 1
            start = pos
            while pos < len(s) and s[pos] not in ",()":
                pos += 1
            bl = float(s[start:pos])
        info[node] = (name, bl)
        children[node] = kids
        return node

    root = parse()
    G = nx.DiGraph()
    G.add_node(root)

    def build(top):
        for c in children[top]:
            G.add_node(c)
            G.add_edge(top, c, weight=info[c][1] or 1.0)
            build(c)

    build(root)
    return G, info
