"""The oracle restatement (oracle/gpe_oracle.cpp) against the golden vectors produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import hashlib

import numpy as np
import pytest

from gnn_pe_b200 import graph_io
from oracle import oracle
from tests.golden_util import CASES, hex_to_f64, load_case


@pytest.fixture(scope="module", params=CASES)
def case(request):
    gold = load_case(request.param)
    g = oracle.OracleGraph.load(gold["data_path"])
    sorted_nodes, membership = graph_io.read_membership(gold["membership_path"], g.V)
    g.enumerate(gold["l"] + 1, sorted_nodes)
    return gold, g, sorted_nodes, membership


def test_label_embedding_kat():
    # SURVEY.md T1/T2 + Appendix C bit patterns (reference gen_vde_x, custom.h:492-511)
    kat = {(2, 0): ["3fda66d132c1bb87", "3fe2cc97669f223d"], (2, 1): ["3fe0892d04810366", "3fdeeda5f6fdf935"],
           (2, 2): ["3fc53759830ca798", "3feab2299f3cd61a"], (2, 36): ["3fbbe36b45ad8a3a", "3fec8392974a4eb9"],
           (4, 0): ["3fc8262fbf4c9287", "3fd13200ed500059", "3fd179549d5b5f2b", "3fd1419295ae5738"],
           (4, 36): ["3fa20109e153bd38", "3fd2686e0feebf57", "3fd52de15f7618c0", "3fd6298f5470b041"]}
    for (e, lab), hexes in kat.items():
        assert oracle.label_embedding(lab, e).tobytes() == hex_to_f64(hexes).tobytes()


def test_graph_meta(case):
    gold, g, _, _ = case
    q0 = gold["queries"][0]
    assert (g.V, g.E, g.labels_count, g.max_degree, g.max_label_freq) == (
        q0["V"], q0["E"], q0["labels_count"], q0["max_degree"], q0["max_label_freq"])


def test_degree_order_matches_membership_file(case):
    gold, g, sorted_nodes, _ = case
    assert np.array_equal(g.degree_order(), sorted_nodes)


def test_path_table_md5(case):
    gold, g, _, membership = case
    assert g.n_rows == gold["n_rows"]
    assert g.paths(0, len(gold["first_rows"])).tolist() == gold["first_rows"]
    assert hashlib.md5(g.all_paths_text().encode()).hexdigest() == gold["all_paths_md5"]
    assert g.rows_per_partition(membership, gold["p"]).tolist() == gold["rows_per_partition"]


def test_closed_form_equals_literal_dfs(case):
    gold, g, sorted_nodes, _ = case
    if gold["n_rows"] > 100000:
        pytest.skip("literal hash-set DFS kept to the small cases")
    closed = g.paths().copy()
    g.enumerate(gold["l"] + 1, sorted_nodes, literal=True)
    assert np.array_equal(g.paths(), closed)
    g.enumerate(gold["l"] + 1, sorted_nodes)


def test_embeddings(case):
    gold, g, _, _ = case
    x, vde = g.embeddings(gold["e"])
    q0 = gold["queries"][0]
    for v, hexes in q0["data_vde_sample"].items():
        assert vde[int(v)].tobytes() == hex_to_f64(hexes).tobytes()
    for lab, hexes in q0["label_x"].items():
        assert oracle.label_embedding(int(lab), gold["e"]).tobytes() == hex_to_f64(hexes).tobytes()


def test_query_plan(case):
    gold, g, _, _ = case
    L, e = gold["l"] + 1, gold["e"]
    for qf, rec in zip(gold["query_paths_files"], gold["queries"]):
        q = oracle.OracleGraph.load(qf)
        for literal in (False, True):
            plan = oracle.query_plan(q, L, e, literal=literal)
            assert plan["n_query_paths"] == rec["n_query_paths"]
            assert len(plan["weight"]) == rec["plan_size"]
            assert plan["vids"].tolist() == [p["vids"] for p in rec["plan"]]
            assert plan["weight"].tolist() == [p["weight"] for p in rec["plan"]]
            for j, p in enumerate(rec["plan"]):
                assert plan["pde"][j].tobytes() == hex_to_f64(p["pde"]).tobytes()


def test_filter_candidates(case):
    gold, g, _, _ = case
    e = gold["e"]
    for qf, rec in zip(gold["query_paths_files"], gold["queries"]):
        q = oracle.OracleGraph.load(qf)
        sets, surv = oracle.filter_candidates(g, q, e)
        assert [s.tolist() for s in sets] == rec["candidates"]
        assert surv.tolist() == [p["survivors"] for p in rec["plan"]]


def test_order_and_answer(case):
    gold, g, _, _ = case
    for qf, rec in zip(gold["query_paths_files"], gold["queries"]):
        if rec["answer"] > 50_000_000:
            continue  # covered by test_big_answer (slow)
        q = oracle.OracleGraph.load(qf)
        cands = [np.array(c, dtype=np.uint32) for c in rec["candidates"]]
        order, pivot = oracle.matching_order(g, q, rec["candidate_counts"])
        assert order.tolist() == rec["order"]
        assert pivot.tolist()[1:] == rec["pivot"][1:]
        limit = rec["limit"] if rec["limit"] is not None else oracle.UINT_MAX
        assert oracle.refine(g, q, cands, limit) == rec["answer"]
        if "main_answer" in rec:
            assert rec["main_answer"] == rec["answer"]


def test_online_end_to_end_quickstart():
    # the reference prints "Answer Number: 45426" on its quick start
    gold = load_case("quickstart")
    g = oracle.OracleGraph.load(gold["data_path"])
    sorted_nodes, _ = graph_io.read_membership(gold["membership_path"], g.V)
    assert g.enumerate(3, sorted_nodes) == 415545
    q = oracle.OracleGraph.load(gold["query_paths_files"][0])
    assert oracle.online(g, q, 2) == 45426
    # the table-free OpenMP variant used as the CPU baseline gives the same answer
    _, vde = g.embeddings(2)
    n, _ = oracle.online_streaming(g, q, 3, 2, sorted_nodes, vde)
    assert n == 45426


def test_matches_dump_is_consistent():
    gold = load_case("uniform300")
    g = oracle.OracleGraph.load(gold["data_path"])
    rec = gold["queries"][0]
    q = oracle.OracleGraph.load(gold["query_paths_files"][0])
    cands = [np.array(c, dtype=np.uint32) for c in rec["candidates"]]
    n, m = oracle.refine(g, q, cands, want_matches=rec["answer"] + 10)
    assert n == rec["answer"] and len(m) == n
    assert len({tuple(r) for r in m.tolist()}) == n
    off, nbr, lab = g.csr()
    qoff, qnbr, qlab = q.csr()
    for r in m[:50]:
        assert len(set(r.tolist())) == q.V
        for u in range(q.V):
            assert lab[r[u]] == qlab[u]
            for w in qnbr[qoff[u]:qoff[u + 1]]:
                assert r[w] in nbr[off[r[u]]:off[r[u] + 1]]


@pytest.mark.slow
def test_big_answer():
    gold = load_case("powerlaw500_e3")
    g = oracle.OracleGraph.load(gold["data_path"])
    rec = gold["queries"][2]
    q = oracle.OracleGraph.load(gold["query_paths_files"][2])
    cands = [np.array(c, dtype=np.uint32) for c in rec["candidates"]]
    assert oracle.refine(g, q, cands) == rec["answer"] == 515204214
