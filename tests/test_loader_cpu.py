"""The `.graph` loader (gpe_host_load_graph; the reference's Static_Graph::loadGraphFromFile, graph.cpp:163-242) on files
the reference would read out of bounds or silently mis-read: an error code every time, never a crash, and well-formed files
unchanged."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from gnn_pe_b200 import gpe, graph_io, synth

GOOD = "t 4 4\nv 0 1 2\nv 1 0 2\nv 2 1 3\nv 3 2 1\ne 0 1\ne 0 2\ne 1 2\ne 2 3\n"


def _load(tmp_path, text):
    p = tmp_path / "g.graph"
    p.write_text(text)
    return gpe.host_load_graph(str(p))


def test_well_formed(tmp_path):
    off, nbr, lab = _load(tmp_path, GOOD)
    assert off.tolist() == [0, 2, 4, 7, 8] and nbr.tolist() == [1, 2, 0, 2, 0, 1, 3, 2] and lab.tolist() == [1, 0, 1, 2]
    off, nbr, lab = _load(tmp_path, "t 0 0\n")
    assert off.tolist() == [0] and len(nbr) == 0 and len(lab) == 0
    off, nbr, lab = _load(tmp_path, "t 2 0\nv 0 5 0\nv 1 6 0\n")
    assert off.tolist() == [0, 0, 0] and lab.tolist() == [5, 6]


@pytest.mark.parametrize("text", [
    "",                                                     # empty file
    "hello\n",                                              # not a graph file
    "v 0 0 0\n",                                            # no header
    GOOD.replace("t 4 4", "t 4 5"),                         # header promises one more edge
    GOOD.replace("t 4 4", "t 4 3"),                         # ... one fewer
    GOOD.replace("v 2 1 3", "v 2 1 2"),                     # declared degree too small
    GOOD.replace("v 2 1 3", "v 2 1 4"),                     # ... too large
    GOOD.replace("e 2 3\n", ""),                            # an edge missing
    GOOD.replace("e 2 3", "e 2 4"),                         # endpoint out of range
    GOOD.replace("v 3 2 1", "v 9 2 1"),                     # vertex id out of range
    GOOD.replace("v 1 0 2\nv 2 1 3\n", "v 2 1 3\nv 1 0 2\n"),  # vertex lines out of order
    GOOD.replace("v 1 0 2\n", ""),                          # a vertex line missing
    GOOD + "e 0\n",                                         # truncated edge line
    GOOD.replace("v 3 2 1", "v 3 2"),                       # truncated vertex line
    "t 4000000000 4000000000\n",                            # sizes that cannot be allocated or do not match
])
def test_malformed_is_an_error(tmp_path, text):
    with pytest.raises(gpe.GpeError):
        _load(tmp_path, text)


def test_round_trip_of_generated_graphs(tmp_path):
    g = synth.chung_lu_graph(500, 2000, 6, gamma=2.5, degree_cap=60, seed=9)
    p = str(tmp_path / "g.graph")
    graph_io.write_graph(p, g)
    off, nbr, lab = gpe.host_load_graph(p)
    assert np.array_equal(off, g.offsets) and np.array_equal(nbr, g.nbrs) and np.array_equal(lab, g.labels)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.tuples(st.sampled_from("tve x"), st.integers(0, 12), st.integers(0, 12), st.integers(0, 12)), max_size=30))
def test_random_token_soup_never_crashes(tmp_path_factory, lines):
    text = "".join(f"{t} {a} {b} {c}\n" if t in "tv" else f"{t} {a} {b}\n" for t, a, b, c in lines)
    p = tmp_path_factory.mktemp("soup") / "g.graph"
    p.write_text(text)
    try:
        off, nbr, lab = gpe.host_load_graph(str(p))
    except gpe.GpeError:
        return
    # accepted: then it is a consistent CSR
    assert off[0] == 0 and (np.diff(off.astype(np.int64)) >= 0).all() and off[-1] == len(nbr)
    assert (nbr < max(len(lab), 1)).all()


# ---- query graphs handed over as arrays (gpe_host_query_plan; the batch calls run the same check) ------------------
TRIANGLE_TAIL = ([0, 2, 4, 7, 8], [1, 2, 0, 2, 0, 1, 3, 2], [1, 0, 1, 2])


def test_query_arrays_well_formed():
    off, nbr, lab = TRIANGLE_TAIL
    plan = gpe.host_query_plan(off, nbr, lab, 3, 2)
    assert len(plan["vids"]) >= 1 and set(plan["vids"].ravel().tolist()) == {0, 1, 2, 3}


@pytest.mark.parametrize("off,nbr", [
    ([1, 2, 4, 7, 8], TRIANGLE_TAIL[1]),                 # offsets do not start at 0
    ([0, 4, 2, 7, 8], TRIANGLE_TAIL[1]),                 # not monotone
    ([0, 2, 4, 7, 4000000000], TRIANGLE_TAIL[1]),        # a degree beyond the query
    (TRIANGLE_TAIL[0], [1, 2, 0, 2, 0, 1, 9, 2]),        # neighbour id out of range
    (TRIANGLE_TAIL[0], [0, 2, 0, 2, 0, 1, 3, 2]),        # self loop
    (TRIANGLE_TAIL[0], [2, 1, 0, 2, 0, 1, 3, 2]),        # adjacency not ascending
    (TRIANGLE_TAIL[0], [1, 1, 0, 2, 0, 1, 3, 2]),        # duplicate edge
    (TRIANGLE_TAIL[0], [1, 2, 0, 2, 0, 1, 3, 1]),        # 3 lists 1, but 1 does not list 3
])
def test_query_arrays_malformed(off, nbr):
    with pytest.raises(gpe.GpeError):
        gpe.host_query_plan(off, nbr, TRIANGLE_TAIL[2], 3, 2)


@pytest.mark.parametrize("off,nbr", [
    ([1, 2, 4, 7, 8], TRIANGLE_TAIL[1]),
    ([0, 4, 2, 7, 8], TRIANGLE_TAIL[1]),
    (TRIANGLE_TAIL[0], [1, 2, 0, 2, 0, 1, 4000000000, 2]),
])
def test_data_arrays_out_of_bounds(off, nbr):
    """Host code that walks a caller's CSR (embeddings, GNN-PGE groups) checks it is memory-safe first."""
    with pytest.raises(gpe.GpeError):
        gpe.host_gen_vde(off, nbr, TRIANGLE_TAIL[2], 2)
    with pytest.raises(gpe.GpeError):
        gpe.host_pge_groups(off, nbr, TRIANGLE_TAIL[2], 2, 2)
