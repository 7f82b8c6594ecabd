"""Large queries (10-16 vertices; 23 of the 56 with more than 16 query paths, where the reference's std::sort is unstable,
SURVEY.md Q4) against the unmodified reference: tests/golden/plans.json, written by tests/golden/make_golden_plans.py.
Holds both the oracle and the product's host plan to the reference's plan order, and the oracle to the reference's
survivors, candidate counts, matching order, pivots and answers under a limit."""
import json
import os

import numpy as np
import pytest

from gnn_pe_b200 import gpe, graph_io
from oracle import oracle
from tests.golden_util import GOLDEN, load_case

SETS = json.load(open(os.path.join(GOLDEN, "plans.json")))


def _query(rec):
    return graph_io.CSRGraph(np.array(rec["offsets"], np.uint32), np.array(rec["nbrs"], np.uint32), np.array(rec["labels"], np.uint32))


@pytest.mark.parametrize("s", SETS, ids=[s["case"] for s in SETS])
def test_plans_of_large_queries(s):
    L, e = s["l"] + 1, s["e"]
    assert sum(r["n_query_paths"] > 16 for r in s["queries"]) >= 5
    for i, rec in enumerate(s["queries"]):
        q = _query(rec)
        oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
        want = np.array(rec["plan"], np.uint32)
        op = oracle.query_plan(oq, L, e)
        assert op["n_query_paths"] == rec["n_query_paths"], i
        assert np.array_equal(op["vids"], want) and op["weight"].tolist() == rec["weights"], i
        hp = gpe.host_query_plan(q.offsets, q.nbrs, q.labels, L, e)
        assert np.array_equal(hp["vids"], want), i
        assert hp["pde"].tobytes() == op["pde"].tobytes(), i


@pytest.mark.parametrize("s", SETS, ids=[s["case"] for s in SETS])
def test_oracle_online_on_large_queries(s):
    gold = load_case(s["case"])
    g = graph_io.read_graph(gold["data_path"])
    sorted_nodes, _ = graph_io.read_membership(gold["membership_path"], g.V)
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    og.enumerate(s["l"] + 1, sorted_nodes)
    for i, rec in enumerate(s["queries"]):
        q = _query(rec)
        oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
        sets, surv = oracle.filter_candidates(og, oq, s["e"])
        assert surv.tolist() == rec["survivors"], i
        assert [len(c) for c in sets] == rec["candidate_counts"], i
        order, pivot = oracle.matching_order(og, oq, rec["candidate_counts"])
        assert order.tolist() == rec["order"] and pivot.tolist()[1:] == rec["pivot"][1:], i
        assert oracle.refine(og, oq, sets, limit=s["limit"]) == rec["answer"], i
