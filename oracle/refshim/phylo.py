"""TEST INFRASTRUCTURE ONLY -- the two Bio.Phylo calls ``TreeData.from_newick`` makes (tree_data.py:187-194):
``Phylo.read(handle, "newick", rooted=True)`` and ``Phylo.to_networkx(tree)``.  Biopython is absent from this image;
restated from its documented behaviour (clades keep the order of the Newick string; ``to_networkx`` adds the root,
then for every child in order: node, edge parent->child with ``weight = branch_length or 1.0``, recursion)."""
import networkx as nx


class Clade:
    def __init__(self, name=None, branch_length=None):
        self.name, self.branch_length, self.clades = name, branch_length, []

    root = property(lambda s: s)

    def __iter__(self):
        return iter(self.clades)

    def __repr__(self):
        return f"Clade(name={self.name!r}, branch_length={self.branch_length!r})"


class Tree:
    def __init__(self, root, rooted):
        self.root, self.rooted = root, rooted

    def prune(self, target=None):
        raise NotImplementedError("outgroup pruning is Bio.Phylo functionality outside the hot path")


def read(handle, fmt, rooted=False):
    assert fmt == "newick"
    s = handle.read().strip().rstrip(";").strip()
    pos = 0

    def parse():
        nonlocal pos
        node = Clade()
        if pos < len(s) and s[pos] == "(":
            pos += 1
            while True:
                node.clades.append(parse())
                if s[pos] == ",":
                    pos += 1
                    continue
                assert s[pos] == ")", f"bad newick at {pos}"
                pos += 1
                break
        start = pos
        while pos < len(s) and s[pos] not in ",():":
            pos += 1
        node.name = s[start:pos].strip() or None
        if pos < len(s) and s[pos] == ":":
            pos += 1
            start = pos
            while pos < len(s) and s[pos] not in ",()":
                pos += 1
            node.branch_length = float(s[start:pos])
        return node

    return Tree(parse(), rooted)


def to_networkx(tree):
    G = nx.DiGraph() if tree.rooted else nx.Graph()
    G.add_node(tree.root)

    def build(top):
        for c in top.clades:
            G.add_node(c)
            G.add_edge(top, c, weight=c.branch_length or 1.0)
            build(c)

    build(tree.root)
    return G
