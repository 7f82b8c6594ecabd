"""The ids-only table layout (GPE_TABLE_IDS: 4L bytes per row, labels / degrees / embeddings gathered by the scan from
packed per-vertex records) against the same golden vectors as the materialised layout: the layout must not change a
single candidate.  This is the layout BASELINE.json's config 4 (l=3, e=4: 160-byte rows) needs at full size."""
import numpy as np
import pytest

from gnn_pe_b200 import gpe, graph_io, synth
from tests.golden_util import CASES, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_ids_only_table_gives_the_golden_results(name):
    gold = load_case(name)
    g = graph_io.read_graph(gold["data_path"])
    sorted_nodes, membership = graph_io.read_membership(gold["membership_path"], g.V)
    L, e = gold["l"] + 1, gold["e"]
    ctx = gpe.GpeContext(0)
    try:
        ctx.set_graph(g.offsets, g.nbrs, g.labels)
        _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, e)
        ctx.set_embeddings(vde)
        n_rows, _ = ctx.enumerate(L, sorted_nodes, membership, gold["p"])
        ctx.set_table_layout(2)
        assert ctx.build_table() == n_rows == gold["n_rows"]
        st = ctx.stats()
        assert st["table_ids_only"] == 1 and st["stored_row_bytes"] == 4 * L and st["row_bytes"] == 8 * L + 8 * L * e
        # the table is the same multiset of rows with the right columns
        rows = ctx.dump_paths()
        vids, labels, degs, pde = ctx.dump_table()
        key = lambda a: a[np.lexsort(a.T[::-1])]
        assert np.array_equal(key(vids), key(rows))
        assert np.array_equal(labels, g.labels[vids]) and np.array_equal(degs, g.degrees[vids])
        assert pde.tobytes() == vde[vids].reshape(len(vids), -1).tobytes()
        # candidate sets and survivor counts, bucketed and streaming
        for flags in (0, gpe.FILTER_NO_PRUNE):
            for qf, rec in zip(gold["query_paths_files"], gold["queries"]):
                qo, qn, ql = gpe.host_load_graph(qf)
                plan = gpe.host_query_plan(qo, qn, ql, L, e)
                sets, surv = ctx.filter(plan, len(ql), flags)
                assert [s.tolist() for s in sets] == rec["candidates"]
                assert surv.tolist() == [p["survivors"] for p in rec["plan"]]
        queries = [graph_io.read_graph(qf) for qf in gold["query_paths_files"]]
        limits = [r["limit"] if r["limit"] is not None else gpe.LIMIT_MAX for r in gold["queries"]]
        assert ctx.query_batch(queries, limits).tolist() == [r["answer"] for r in gold["queries"]]
        # back to materialised rows on the same context
        ctx.set_table_layout(1)
        ctx.build_table()
        assert ctx.stats()["table_ids_only"] == 0
        assert ctx.query_batch(queries, limits).tolist() == [r["answer"] for r in gold["queries"]]
    finally:
        ctx.close()


def test_ids_only_equals_rows_on_a_config4_shaped_batch():
    """l=3, e=4, dense 12-vertex queries (config 4 scaled down): both layouts, same candidate sets and answers."""
    g = synth.chung_lu_graph(3000, 12000, 6, gamma=2.8, degree_cap=40, seed=43)
    queries = synth.query_batch(g, 8, 12, seed=44)
    sorted_nodes = graph_io.degree_order(g)
    _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, 4)
    out = {}
    for layout in (1, 2):
        ctx = gpe.GpeContext(0)
        try:
            ctx.set_graph(g.offsets, g.nbrs, g.labels)
            ctx.set_embeddings(vde)
            ctx.enumerate(4, sorted_nodes, graph_io.block_membership(g.V, 4), 4)
            ctx.set_table_layout(layout)
            ctx.build_table()
            ans = ctx.query_batch(queries, [2_000_000] * len(queries)).tolist()
            off, cand = ctx.batch_get_candidates()
            out[layout] = (ans, off.tolist(), cand.tolist())
        finally:
            ctx.close()
    assert out[1] == out[2] and sum(1 for a in out[1][0] if a > 0) >= 3
