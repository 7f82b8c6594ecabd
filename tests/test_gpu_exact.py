"""Exact mode (GPE_FILTER_BOTH_ORIENTATIONS): every plan path compared with both orientations of every
stored row.  The candidate sets are then complete and the answer is the true number of embeddings -- label equal, degree
>=, every query edge on a data edge, injective -- which tests/bigcount.py counts without any filter."""
import numpy as np
import pytest

from gnn_pe_b200 import engine, gpe, graph_io, synth
from tests import bigcount
from tests.golden_util import load_case

pytestmark = pytest.mark.gpu


def _truth(g, q):
    start = 0
    deg = g.degrees
    pool = np.nonzero((g.labels == q.labels[start]) & (deg >= q.degrees[start]))[0]
    return bigcount.exact_count(g, q, start, pool)


def test_quickstart_true_count():
    gold = load_case("quickstart")
    g = graph_io.read_graph(gold["data_path"])
    q = graph_io.read_graph(gold["query_paths_files"][0])
    sorted_nodes, membership = graph_io.read_membership(gold["membership_path"], g.V)
    eng = engine.Engine(0)
    try:
        eng.offline(g, l=2, e=2, p=5, sorted_nodes=sorted_nodes, membership=membership)
        assert int(eng.ctx.query_batch([q])[0]) == 45426                                      # the reference's answer
        assert int(eng.ctx.query_batch([q], flags=gpe.FILTER_BOTH_ORIENTATIONS)[0]) == 221832
        assert _truth(g, q) == 221832
    finally:
        eng.close()


@pytest.mark.parametrize("seed,nl", [(3, 4), (4, 7)])
def test_random_queries_true_count(seed, nl):
    g = synth.chung_lu_graph(1500, 7000, nl, gamma=2.6, degree_cap=50, seed=seed)
    queries = synth.query_batch(g, 14, (4, 8), seed=seed + 10, mixed=True)
    eng = engine.Engine(0)
    try:
        eng.offline(g, l=2, e=2, p=3)
        ref = eng.ctx.query_batch(queries).tolist()
        exact = eng.ctx.query_batch(queries, flags=gpe.FILTER_BOTH_ORIENTATIONS).tolist()
        truth = [_truth(g, q) for q in queries]
        assert exact == [min(t, gpe.LIMIT_MAX) for t in truth]
        assert all(r <= x for r, x in zip(ref, exact)) and ref != exact  # the reference's rule under-counts somewhere
        # streaming (no bucket pruning) scan, same flag
        assert eng.ctx.query_batch(queries, flags=gpe.FILTER_BOTH_ORIENTATIONS | gpe.FILTER_NO_PRUNE).tolist() == exact
    finally:
        eng.close()


@pytest.mark.parametrize("seed,nl,l,e", [(6, 4, 2, 2), (7, 3, 3, 2)])
def test_exact_candidate_sets_match_oracle(seed, nl, l, e):
    """gpe_filter with the flag: candidate lists and survivors equal to the oracle's exact mode (itself pinned against
    the filter-free count in tests/test_exact_oracle_cpu.py), bucketed and streaming scans."""
    from oracle import oracle
    g = synth.chung_lu_graph(600, 2600, nl, gamma=2.6, degree_cap=40, seed=seed)
    eng = engine.Engine(0)
    try:
        eng.offline(g, l=l, e=e, p=3)
        og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
        og.enumerate(l + 1, graph_io.degree_order(g))
        for q in synth.query_batch(g, 6, (4, 8), seed=seed + 30, mixed=True):
            oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
            sets, surv = oracle.filter_candidates(og, oq, e, both_orientations=True)
            plan = gpe.host_query_plan(q.offsets, q.nbrs, q.labels, l + 1, e)
            for flags in (gpe.FILTER_BOTH_ORIENTATIONS, gpe.FILTER_BOTH_ORIENTATIONS | gpe.FILTER_NO_PRUNE):
                gsets, gsurv = eng.ctx.filter(plan, q.V, flags)
                assert [s.tolist() for s in gsets] == [s.tolist() for s in sets]
                assert gsurv.tolist() == surv.tolist()
    finally:
        eng.close()
