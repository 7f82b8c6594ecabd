"""tests/bigcount.py (the arbitrary-precision counter the overflow tests rely on) against the oracle, which
enumerates like the reference: equal wherever the oracle can finish, i.e. small counts and limited runs."""
import numpy as np
import pytest

from gnn_pe_b200 import graph_io, synth
from oracle import oracle
from tests import bigcount, overflow_cases


@pytest.mark.parametrize("name", list(overflow_cases.CASES))
def test_exact_count_matches_oracle_under_limits(name):
    g, q = overflow_cases.CASES[name]()
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
    og.enumerate(3, graph_io.degree_order(g))
    for limit in (1, 1000, 200_000):
        want, total = bigcount.reference_answer(g, q, 2, 2, limit)
        assert oracle.online(og, oq, 2, limit) == want, (name, limit, total)
    if total < 10_000_000:
        assert oracle.online(og, oq, 2) == total


def test_exact_count_matches_oracle_on_random_queries():
    g = synth.uniform_graph(400, 2400, 5, seed=11)
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    og.enumerate(3, graph_io.degree_order(g))
    for i, q in enumerate(synth.query_batch(g, 12, (4, 9), seed=5, mixed=True)):
        oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
        want, total = bigcount.reference_answer(g, q, 2, 2, oracle.UINT_MAX)
        assert oracle.online(og, oq, 2) == want == total, i


def test_the_big_cases_are_big():
    _, t1 = bigcount.reference_answer(*overflow_cases.CASES["star_2^72"](), 2, 2, 1)
    _, t2 = bigcount.reference_answer(*overflow_cases.CASES["pair_3x2^64"](), 2, 2, 1)
    _, t3 = bigcount.reference_answer(*overflow_cases.CASES["pair_1x2^64"](), 2, 2, 1)
    _, t4 = bigcount.reference_answer(*overflow_cases.CASES["minuend_saturated"](), 2, 2, 1)
    assert t1 % (1 << 64) == 0 and t1 > 0          # a wrapping 64-bit product reports 0
    assert t2 == 3 << 64 and t3 == 1 << 64         # so do these sums of products 2^32 x 2^32
    assert t4 == 1
