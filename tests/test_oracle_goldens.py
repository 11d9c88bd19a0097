"""Pins the oracle (and the product's host-side mirror) to every golden value the reference holds for
this path: the doctests of tree_data.py and util.py.  CPU only."""
import networkx as nx
import numpy as np
import pytest

import oracle.tree_data as otd
from oracle.util import order_events
from vbsky_b200.tree_data import TreeData

IMPLS = [otd.TreeData, TreeData]


@pytest.mark.parametrize("TD", IMPLS)
def test_lower_sampling_times_doctests(TD):
    # tree_data.py:88-100
    assert TD.from_newick("(a,b)")[0].lower_sampling_times.tolist() == [0.0, 0.0, 0.0]
    assert TD.from_newick("(a:1,b:2)")[0].lower_sampling_times.tolist() == [1.0, 0.0, 1.0]
    assert TD.from_newick("((a:1,b:2):3,c:4)")[0].lower_sampling_times.tolist() == [1.0, 0.0, 1.0, 1.0, 1.0]


@pytest.mark.parametrize("TD", IMPLS)
def test_from_newick_doctests(TD):
    # tree_data.py:164-183
    td, nm = TD.from_newick("((human:1,chimp:2):1,gorilla:2.5);")
    assert (nm["human"], nm["chimp"], nm["gorilla"]) == (0, 1, 2)
    assert np.array_equal(td.sample_times, [1.0, 0.0, 0.5])
    assert td.n == 3
    assert np.array_equal(TD.from_newick("((human,chimp),gorilla)")[0].sample_times, [0.0, 0.0, 1.0])
    assert TD.from_newick("(A,B,C,D)")[0].to_newick() == "(((3,2),1),0)"  # '(((D,C),B),A)' by index


@pytest.mark.parametrize("TD", IMPLS)
def test_to_newick_doctests(TD):
    # tree_data.py:122-131
    t = TD.from_newick("(A:1,B:2)")[0]
    assert t.to_newick() == "(1,0)"
    assert t.to_newick([1, 2], node_mapping={0: "A", 1: "B"}) == "(B:2,A:1)"


@pytest.mark.parametrize("TD", IMPLS)
def test_height_transform_doctests(TD):
    # tree_data.py:260-269 and :317-321.  The docstrings were written for contemporaneous tips
    # (sample_times == 0); with the current from_newick an unlabelled "((humans,chimp),gorilla)" has
    # sample_times [0,0,1], so the golden numbers are checked on the ultrametric version of that tree.
    td = TD.from_newick("((humans:1,chimp:1):1,gorilla:2);")[0]
    assert np.array_equal(td.sample_times, [0.0, 0.0, 0.0])
    np.testing.assert_allclose(td.height_transform(3.0, [0.5]), [0.0, 0.0, 0.0, 1.5, 3.0], rtol=0, atol=1e-15)
    heights = np.array([0.0, 0.0, 0.0, 2.0, 5.0])
    root_height, proportions = td.inverse_height_transform(heights)
    assert root_height == 5.0
    np.testing.assert_allclose(proportions, [0.4], rtol=1e-15)
    assert td.height_transform(root_height, proportions).tolist() == heights.tolist()
    # as coded (tree_data.py:351-357) heterochronous tips shift the root by max(sample_times)
    td2 = TD.from_newick("((humans,chimp),gorilla);")[0]
    np.testing.assert_allclose(td2.height_transform(3.0, [0.5]), [0.0, 0.0, 1.0, 2.0, 4.0], rtol=0, atol=1e-15)


@pytest.mark.parametrize("TD", IMPLS)
def test_from_nx_doctest(TD):
    # tree_data.py:433-461
    G = nx.DiGraph()
    G.add_edges_from([("root", "leaf1"), ("root", "leaf2")])
    td, nm = TD.from_nx(G)
    assert td.n == 2
    assert nm == {"leaf1": 0, "leaf2": 1, "root": 2}
    assert td.postorder.tolist() == [0, 1, 2] and td.preorder.tolist() == [2, 1, 0]
    assert td.sample_times.tolist() == [0.0, 0.0]  # the doctest's `array([0., 1.])` line contradicts unit default lengths
    assert td.child_parent.tolist() == [2, 2, -1]
    assert td.parent_child.tolist() == [[-1, -1], [-1, -1], [1, 0]]


def test_order_events_doctest():
    # util.py:41-62
    ev = order_events([0, 0.5, 1.0], [0.0, 0.0, 0.1, 0.7, 0.7, 0.9], [True, True, False, True, True, False]).tolist()
    assert len(ev) == 9
    assert sorted(ev[:3]) == [[0.0, 0.0], [0.0, 1.0], [0.0, 2.0]]
    assert ev[3] == [0.1, 1.0] and ev[4] == [0.5, 1.0]
    assert ev[5:7] == [[0.7, 2.0], [0.7, 3.0]]
    assert ev[7] == [0.9, 2.0] and ev[8] == [1.0, 2.0]
