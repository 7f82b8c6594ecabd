"""Large queries (10-16 vertices, tests/golden/plans.json: the unmodified reference's survivors, candidate counts, matching
order and answers under a limit) through the C ABI: gpe_filter per query, then the whole batch in one gpe_query_batch."""
import json
import os

import numpy as np
import pytest

from gnn_pe_b200 import engine, gpe, graph_io
from tests.golden_util import GOLDEN, load_case

pytestmark = pytest.mark.gpu
SETS = json.load(open(os.path.join(GOLDEN, "plans.json")))


@pytest.mark.parametrize("s", SETS, ids=[s["case"] for s in SETS])
def test_large_queries_against_reference(s):
    gold = load_case(s["case"])
    g = graph_io.read_graph(gold["data_path"])
    sorted_nodes, membership = graph_io.read_membership(gold["membership_path"], g.V)
    queries = [graph_io.CSRGraph(np.array(r["offsets"], np.uint32), np.array(r["nbrs"], np.uint32),
                                 np.array(r["labels"], np.uint32)) for r in s["queries"]]
    eng = engine.Engine(0)
    try:
        eng.offline(g, l=s["l"], e=s["e"], p=s["p"], sorted_nodes=sorted_nodes, membership=membership)
        for i, (q, rec) in enumerate(zip(queries, s["queries"])):
            plan = gpe.host_query_plan(q.offsets, q.nbrs, q.labels, s["l"] + 1, s["e"])
            sets, surv = eng.ctx.filter(plan, q.V)
            assert surv.tolist() == rec["survivors"], i
            assert [len(c) for c in sets] == rec["candidate_counts"], i
            res = eng.ctx.refine(q.offsets, q.nbrs, q.labels, sets, limit=s["limit"])
            assert res["order"].tolist() == rec["order"] and res["pivot"].tolist()[1:] == rec["pivot"][1:], i
            assert res["n_matches"] == rec["answer"], i
        ans = eng.ctx.query_batch(queries, [s["limit"]] * len(queries))
        assert ans.tolist() == [r["answer"] for r in s["queries"]]
    finally:
        eng.close()
