"""Counts beyond 64 bits and the answer limit: the reference counts one embedding at a time and stops at the
limit (custom.h:846-855), so its answer is min(total, limit) with limit <= UINT_MAX whatever the total.  The GPU join
multiplies table entries instead; these cases have totals of 2^72, 3 x 2^64 and 2^64 (a wrapping product reports 0)
and one whose answer is 1 while an intermediate table sum exceeds 2^62.  Expected answers: tests/bigcount.py
(arbitrary precision), itself held to the oracle in tests/test_bigcount_cpu.py."""
import numpy as np
import pytest

from gnn_pe_b200 import engine, gpe, graph_io
from tests import bigcount, overflow_cases
from tests.golden_util import load_case

pytestmark = pytest.mark.gpu

LIMITS = [gpe.LIMIT_MAX, 10**6, 1, 0]


@pytest.mark.parametrize("name", list(overflow_cases.CASES))
def test_answer_is_min_of_total_and_limit(name):
    g, q = overflow_cases.CASES[name]()
    eng = engine.Engine(0)
    try:
        eng.offline(g, l=2, e=2, p=2)
        want = [bigcount.reference_answer(g, q, 2, 2, lim)[0] for lim in LIMITS]
        total = bigcount.reference_answer(g, q, 2, 2, 1)[1]
        got = [eng.online(q, lim) for lim in LIMITS]
        assert got == want, (name, total)
        # the same query four times in one batch, one limit each
        assert eng.online_batch([q] * len(LIMITS), LIMITS).tolist() == want
        # raw (unclamped) totals: exact below 2^44, saturated beyond -- never wrapped
        eng.ctx.batch_upload([q], [gpe.LIMIT_MAX])
        eng.ctx.batch_filter()
        eng.ctx.batch_join(0, 1)
        raw = int(eng.ctx.batch_download()[0])
        assert raw == total if total < (1 << 44) else (1 << 44) <= raw <= (1 << 48)
        if name.startswith("minuend_saturated"):
            assert eng.ctx.stats()["join_reruns"] >= 1  # the weighted count met a saturated sum and was walked instead
    finally:
        eng.close()


def test_small_limit_stops_the_join_early():
    """`-n N` ends the enumeration (custom.h:851-854): far fewer DFS steps than the unlimited run."""
    from gnn_pe_b200 import synth
    g = synth.uniform_graph(3000, 36000, 2, seed=7)
    eng = engine.Engine(0)
    try:
        eng.offline(g, l=2, e=2, p=4)
        best, full, total = None, 0, 0
        for q in synth.query_batch(g, 6, 7, seed=8):  # dense 7-vertex queries over 2 labels: cores with repeated labels
            n = eng.online(q)
            steps = eng.ctx.stats()["join_steps"]
            if steps > full:
                best, full, total = q, steps, n
        assert full > 200_000 and total > 1000, (full, total)  # the case must be heavy for the comparison to mean something
        assert eng.online(best, 10) == 10
        limited = eng.ctx.stats()["join_steps"]
        assert limited * 4 < full, (limited, full)
        # copies with different limits in one batch do not disturb each other
        assert eng.online_batch([best, best, best], [5, gpe.LIMIT_MAX, 1000]).tolist() == [5, total, 1000]
    finally:
        eng.close()
