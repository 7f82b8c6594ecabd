"""gnn_pe_b200.prepare -- the reference's gnnpe.py step (folders + membership.txt) without METIS."""
import os

import numpy as np
import pytest

from gnn_pe_b200 import graph_io, prepare, synth
from tests.golden_util import load_case


@pytest.mark.parametrize("how", ["rcm", "block", "auto"])
def test_layout_and_membership(tmp_path, how):
    gold = load_case("quickstart")
    d = str(tmp_path) + "/"
    assert prepare.main(["--f", d, "--d", gold["data_path"], "--p", "5", "--partitioner", how]) == 0
    for i in range(5):
        assert os.path.isdir(d + f"gnn-pe/partitions/partition-{i}")
    g = graph_io.read_graph(gold["data_path"])
    sorted_nodes, membership = graph_io.read_membership(d + "gnn-pe/membership.txt", g.V)
    # line order = the reference's: ascending degree, stable (the committed membership.txt of the golden case has it too)
    want_sorted, _ = graph_io.read_membership(gold["membership_path"], g.V)
    assert np.array_equal(sorted_nodes, want_sorted)
    sizes = np.bincount(membership, minlength=5)
    assert membership.max() < 5 and sizes.min() >= g.V // 5 and sizes.max() <= g.V // 5 + 1
    # a second run replaces the tree (gnnpe.py:57)
    open(d + "gnn-pe/stale.txt", "w").write("x")
    prepare.main(["--f", d, "--d", gold["data_path"], "--p", "3", "--partitioner", how])
    assert not os.path.exists(d + "gnn-pe/stale.txt") and not os.path.isdir(d + "gnn-pe/partitions/partition-3")


def test_rcm_keeps_neighbours_together(tmp_path):
    # a ring of cliques with shuffled vertex ids: id blocks cut almost every edge, an order-based cut almost none
    rng = np.random.default_rng(5)
    k, n = 8, 40
    perm = rng.permutation(k * n)
    edges = [(perm[c * k + i], perm[c * k + j]) for c in range(n) for i in range(k) for j in range(i + 1, k)]
    edges += [(perm[c * k], perm[((c + 1) % n) * k + 1]) for c in range(n)]
    g = graph_io.csr_from_edges(k * n, np.array(edges), rng.integers(0, 4, k * n).astype(np.uint32))
    cut_rcm = prepare.edge_cut(g, prepare.partition(g, 4, "rcm"))
    cut_block = prepare.edge_cut(g, prepare.partition(g, 4, "block"))
    assert cut_rcm * 10 < cut_block


def test_pge_variant_and_errors(tmp_path):
    g = synth.uniform_graph(50, 120, 3, seed=1)
    p = str(tmp_path / "g.graph")
    graph_io.write_graph(p, g)
    assert prepare.main(["--f", str(tmp_path), "--d", p, "--p", "2", "--variant", "pge", "--partitioner", "block"]) == 0
    assert os.path.isfile(str(tmp_path / "gnn-pge" / "membership.txt"))
    with pytest.raises(ValueError):
        prepare.partition(g, 0)


REF_TEST = "/root/reference/Test/"


@pytest.mark.skipif(not os.path.exists(REF_TEST + "data_graph.gpickle.gz"), reason="the reference tree is not on this machine")
def test_reads_the_reference_pickle(tmp_path):
    """The reference's script takes the networkx pickle of the data graph; the same topology comes out of either file."""
    pytest.importorskip("networkx")
    a, b = str(tmp_path / "a") + "/", str(tmp_path / "b") + "/"
    prepare.main(["--f", a, "--d", REF_TEST + "data_graph.gpickle.gz", "--p", "5", "--partitioner", "rcm"])
    prepare.main(["--f", b, "--d", REF_TEST + "data_graph.graph", "--p", "5", "--partitioner", "rcm"])
    assert open(a + "gnn-pe/membership.txt").read() == open(b + "gnn-pe/membership.txt").read()
