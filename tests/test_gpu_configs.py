"""BASELINE.json's other configurations as parity cases: scaled copies the oracle finishes in seconds, and the
full-size config 2 through size-independent properties (closed-form row count, path validity, pruning and batch
composition invariance, the limit rule) plus one query checked against the oracle's all-pairs compare.

  config 4: longer paths (l=3, e=4) and dense 12-vertex queries         -> test_config4_l3_e4_dense_queries
  config 5: a mixed batch of sparse / dense queries of 4-16 vertices     -> test_config5_mixed_batch
  config 2 at full size (1 M vertices / 10 M edges / 514.9 M table rows) -> test_config2_full_size_properties
"""
import numpy as np
import pytest

from gnn_pe_b200 import gpe, graph_io, synth

pytestmark = pytest.mark.gpu


def _engine(g, l, e, p=8):
    ctx = gpe.GpeContext(0)
    ctx.set_graph(g.offsets, g.nbrs, g.labels)
    _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, e)
    ctx.set_embeddings(vde)
    sorted_nodes = graph_io.degree_order(g)
    n_rows, rows_pp = ctx.enumerate(l + 1, sorted_nodes, graph_io.block_membership(g.V, p), p)
    ctx.build_table()
    return ctx, vde, sorted_nodes, n_rows, rows_pp


def test_config4_l3_e4_dense_queries():
    """Config 4 scaled down: l=3 (4-vertex paths, 160-byte rows), e=4, induced 12-vertex queries."""
    from oracle import oracle
    g = synth.chung_lu_graph(4000, 16000, 6, gamma=2.8, degree_cap=40, seed=41)
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    ctx, vde, sorted_nodes, n_rows, _ = _engine(g, 3, 4, p=4)
    assert n_rows == og.enumerate(4, sorted_nodes)
    assert ctx.stats()["row_bytes"] == 160
    queries = synth.query_batch(g, 10, 12, seed=42)
    limit = 5_000_000
    expect = []
    for q in queries:
        oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
        n, _ = oracle.online_streaming(og, oq, 4, 4, sorted_nodes, vde, limit=limit, threads=4)
        expect.append(n)
    assert ctx.query_batch(queries, [limit] * len(queries)).tolist() == expect
    assert sum(1 for x in expect if x > 0) >= 5
    # streaming (no pruning) scan gives the same candidate sets as the bucketed one
    q = queries[0]
    plan = gpe.host_query_plan(q.offsets, q.nbrs, q.labels, 4, 4)
    a, sa = ctx.filter(plan, q.V)
    b, sb = ctx.filter(plan, q.V, gpe.FILTER_NO_PRUNE)
    assert [x.tolist() for x in a] == [x.tolist() for x in b] and sa.tolist() == sb.tolist()
    ctx.close()


def test_config5_mixed_batch():
    """Config 5 scaled down: one batch of sparse (tree) and dense (induced) queries of 4-16 vertices, some with an
    answer limit; the 9-16-vertex queries run the wider stack geometries of the join."""
    from oracle import oracle
    g = synth.chung_lu_graph(6000, 30000, 12, gamma=2.6, degree_cap=60, seed=51)
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    ctx, vde, sorted_nodes, n_rows, _ = _engine(g, 2, 2)
    assert og.enumerate(3, sorted_nodes) == n_rows
    queries = synth.query_batch(g, 48, (4, 16), seed=52, mixed=True)
    assert max(q.V for q in queries) > 8 and min(q.V for q in queries) <= 6
    limits = [gpe.LIMIT_MAX if i % 3 else 1000 for i in range(len(queries))]
    cap = 20_000_000  # keeps the oracle's enumeration bounded; both sides apply the same rule
    limits = [min(l, cap) for l in limits]
    expect = [oracle.online(og, oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels), 2, l)
              for q, l in zip(queries, limits)]
    ans = ctx.query_batch(queries, limits).tolist()
    bad = [(i, queries[i].V, a, e) for i, (a, e) in enumerate(zip(ans, expect)) if a != e]
    assert not bad, bad
    assert sum(1 for e in expect if e > 0) >= len(expect) // 2
    # the same queries in a different batch composition (reversed, and only the large ones)
    assert ctx.query_batch(queries[::-1], limits[::-1]).tolist() == expect[::-1]
    big = [i for i, q in enumerate(queries) if q.V > 8]
    assert ctx.query_batch([queries[i] for i in big], [limits[i] for i in big]).tolist() == [expect[i] for i in big]
    ctx.close()


def test_config2_full_size_properties():
    """BASELINE.json configs[1] at full size.  Nothing here needs the oracle to enumerate 0.5 G rows."""
    import bench
    from oracle import oracle
    w, g, queries = bench.load_workload("config2")
    ctx, vde, sorted_nodes, n_rows, rows_pp = _engine(g, w["l"], w["e"], w["p"])
    # (1) row count: closed form sum C(deg, 2); partitions add up; start rows are a prefix sum
    assert n_rows == synth.table_rows_l2(g) == int(rows_pp.sum())
    sr = ctx.start_rows()
    assert sr[0] == 0 and sr[-1] == n_rows and np.all(np.diff(sr.astype(np.int64)) >= 0)
    d = g.degrees.astype(np.int64)[sorted_nodes]  # rows of a start vertex a: pairs (b, c), b in N(a), c in N(b) later in
    assert int(np.diff(sr.astype(np.int64)).max()) <= int((d * g.degrees.max()).max())  # the order: bounded by deg(a) x max deg
    # (2) windows of all_paths.txt against the closed form of the reference's dfs + hash-set dedup (custom.h:66-92,
    #     SURVEY.md 3.1): start vertices in membership order; for a: b over N(a) ascending, c over N(b) ascending,
    #     kept iff c comes later than a in that order (the reverse orientation is what the set would reject)
    rank = np.empty(g.V, dtype=np.int64)
    rank[sorted_nodes] = np.arange(g.V)
    off, nbr = g.offsets.astype(np.int64), g.nbrs.astype(np.int64)

    def rows_of(r):
        a = int(sorted_nodes[r])
        out = []
        for b in nbr[off[a]:off[a + 1]]:
            cs = nbr[off[b]:off[b + 1]]
            cs = cs[rank[cs] > r]
            out.append(np.stack([np.full(len(cs), a), np.full(len(cs), b), cs], axis=1))
        return np.concatenate(out) if out else np.zeros((0, 3), dtype=np.int64)

    sr64 = sr.astype(np.int64)
    for first in (0, n_rows // 3, n_rows - 20_000):
        got = ctx.dump_paths(first, 20_000).astype(np.int64)
        r0 = int(np.searchsorted(sr64, first, side="right") - 1)
        r1 = int(np.searchsorted(sr64, first + 20_000 - 1, side="right") - 1)
        want = np.concatenate([rows_of(r) for r in range(r0, r1 + 1)])
        lo = first - int(sr64[r0])
        assert np.array_equal(got, want[lo:lo + 20_000])
    # (3) pruning never changes the result: bucketed and streaming scans give identical candidate sets
    L, e = w["l"] + 1, w["e"]
    q = queries[7]
    plan = gpe.host_query_plan(q.offsets, q.nbrs, q.labels, L, e)
    a, sa = ctx.filter(plan, q.V)
    b, sb = ctx.filter(plan, q.V, gpe.FILTER_NO_PRUNE)
    assert [x.tolist() for x in a] == [x.tolist() for x in b] and sa.tolist() == sb.tolist()
    st = ctx.stats()
    assert st["scan_rows"] >= n_rows  # the streaming pass looked at every row
    # (4) batch composition invariance + the limit rule min(N, total)
    full = ctx.query_batch(queries)
    assert int((full > 0).sum()) == len(queries)  # random-walk queries have at least the walk itself
    sub = [3, 17, 42, 62, 99]
    assert ctx.query_batch([queries[i] for i in sub]).tolist() == [int(full[i]) for i in sub]
    lim = [1, 1000, 10**6, 10**7, gpe.LIMIT_MAX]
    got = ctx.query_batch([queries[i] for i in sub], lim).tolist()
    assert got == [min(int(full[i]), l) for i, l in zip(sub, lim)]
    # (5) one query against the oracle's all-pairs leaf compare + the reference's refinement
    qi = int(np.argmin(full))
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    oq = oracle.OracleGraph.from_csr(queries[qi].offsets, queries[qi].nbrs, queries[qi].labels)
    n, _ = oracle.online_streaming(og, oq, L, e, sorted_nodes, vde)
    assert n == int(full[qi])
    ctx.close()


@pytest.mark.parametrize("name", ["config3_small", "config5_small", "small"])
def test_bench_workloads_match_the_oracle_fixture(name):
    """Down-scaled copies of bench.py's config 3 / 5 workloads (same generator, labels and query mix) and of config 2:
    the sampled queries of tests/golden/config_answers.json (oracle.online_streaming on the CPU) through the batch path."""
    import json
    import os
    import bench
    from tests.golden_util import ROOT
    rec = json.load(open(os.path.join(ROOT, "tests", "golden", "config_answers.json")))[name]
    w, g, queries = bench.load_workload(name)
    assert (g.V, g.E, len(queries)) == (rec["V"], rec["E"], rec["n_queries"])  # the generators are deterministic
    ctx, vde, sorted_nodes, n_rows, _ = _engine(g, w["l"], w["e"], p=w["p"])
    try:
        ans = ctx.query_batch(queries)
        got = {i: int(ans[int(i)]) for i in rec["answers"]}
        assert got == {i: min(v, gpe.LIMIT_MAX) for i, v in rec["answers"].items()}
    finally:
        ctx.close()
