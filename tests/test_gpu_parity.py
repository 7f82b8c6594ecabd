"""Parity tests proper: the CUDA path (through the C ABI) against the golden vectors of the unmodified
reference and against the oracle on seeded inputs.  Bit-exact everywhere: integer ids, candidate sets and
counts; the FP64 compares consume identical values so no tolerance is involved."""
import hashlib

import numpy as np
import pytest

from gnn_pe_b200 import gpe, graph_io, sharding, synth
from tests.golden_util import CASES, load_case

pytestmark = pytest.mark.gpu


def _ctx_for(gold):
    g = graph_io.read_graph(gold["data_path"])
    sorted_nodes, membership = graph_io.read_membership(gold["membership_path"], g.V)
    ctx = gpe.GpeContext(0)
    ctx.set_graph(g.offsets, g.nbrs, g.labels)
    _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, gold["e"])
    ctx.set_embeddings(vde)
    n_rows, rows_pp = ctx.enumerate(gold["l"] + 1, sorted_nodes, membership, gold["p"])
    return ctx, g, vde, sorted_nodes, membership, n_rows, rows_pp


@pytest.fixture(scope="module", params=CASES)
def case(request):
    gold = load_case(request.param)
    ctx, g, vde, sorted_nodes, membership, n_rows, rows_pp = _ctx_for(gold)
    ctx.build_table()
    yield gold, ctx, g, vde, sorted_nodes, membership, n_rows, rows_pp
    ctx.close()


def _text(rows):
    return f"{len(rows)}\n" + "".join(" ".join(map(str, r)) + " \n" for r in rows.tolist())


def test_enumerate_matches_reference_all_paths(case):
    gold, ctx, g, vde, sorted_nodes, membership, n_rows, rows_pp = case
    assert n_rows == gold["n_rows"]
    assert rows_pp.tolist() == gold["rows_per_partition"]
    rows = ctx.dump_paths()
    assert rows[: len(gold["first_rows"])].tolist() == gold["first_rows"]
    assert hashlib.md5(_text(rows).encode()).hexdigest() == gold["all_paths_md5"]
    # windows that start and end inside a start vertex's block
    for first, n in [(1, 1), (n_rows // 3, 1000), (n_rows - 7, 7)]:
        assert np.array_equal(ctx.dump_paths(first, n), rows[first:first + n])
    # partition_paths.txt content (T6): ids of the paths whose first vertex is in the partition
    start = ctx.start_rows()
    assert start[-1] == n_rows
    firsts_member = membership[rows[:, 0]]
    for i in range(gold["p"]):
        assert int((firsts_member == i).sum()) == gold["rows_per_partition"][i]


def test_table_is_the_same_multiset_with_right_columns(case):
    gold, ctx, g, vde, *_ = case
    rows = ctx.dump_paths()
    vids, labels, degs, pde = ctx.dump_table()
    assert len(vids) == gold["n_rows"]
    key = lambda a: a[np.lexsort(a.T[::-1])]
    assert np.array_equal(key(vids), key(rows))
    deg = g.degrees
    assert np.array_equal(labels, g.labels[vids])
    assert np.array_equal(degs, deg[vids])
    assert pde.tobytes() == vde[vids].reshape(len(vids), -1).tobytes()


@pytest.mark.parametrize("flags", [0, gpe.FILTER_NO_PRUNE])
def test_filter_candidates_and_survivors(case, flags):
    gold, ctx, *_ = case
    L, e = gold["l"] + 1, gold["e"]
    for qf, rec in zip(gold["query_paths_files"], gold["queries"]):
        qo, qn, ql = gpe.host_load_graph(qf)
        plan = gpe.host_query_plan(qo, qn, ql, L, e)
        sets, surv = ctx.filter(plan, len(ql), flags)
        assert [s.tolist() for s in sets] == rec["candidates"]
        assert surv.tolist() == [p["survivors"] for p in rec["plan"]]
    st = ctx.stats()
    if flags:
        assert st["scan_items"] == st["scan_items_unpruned"]


def test_refine_order_and_answer(case):
    gold, ctx, *_ = case
    for qf, rec in zip(gold["query_paths_files"], gold["queries"]):
        qo, qn, ql = gpe.host_load_graph(qf)
        cands = [np.array(c, dtype=np.uint32) for c in rec["candidates"]]
        limit = rec["limit"] if rec["limit"] is not None else gpe.LIMIT_MAX
        res = ctx.refine(qo, qn, ql, cands, limit)
        assert res["order"].tolist() == rec["order"]
        assert res["pivot"].tolist()[1:] == rec["pivot"][1:]
        assert res["n_matches"] == rec["answer"]


def test_query_batch_end_to_end(case):
    gold, ctx, *_ = case
    queries = [graph_io.read_graph(qf) for qf in gold["query_paths_files"]]
    limits = [r["limit"] if r["limit"] is not None else gpe.LIMIT_MAX for r in gold["queries"]]
    ans = ctx.query_batch(queries, limits)
    assert ans.tolist() == [r["answer"] for r in gold["queries"]]
    # same batch, streaming (unpruned) scan
    ans2 = ctx.query_batch(queries, limits, flags=gpe.FILTER_NO_PRUNE)
    assert ans2.tolist() == ans.tolist()
    # candidate sets of the batch equal the per-query golden sets
    off, cand = ctx.batch_get_candidates()
    slot = 0
    for r in gold["queries"]:
        for c in r["candidates"]:
            assert cand[int(off[slot]):int(off[slot + 1])].tolist() == c
            slot += 1


def test_match_set_equals_oracle():
    from oracle import oracle
    gold = load_case("uniform300")
    ctx, g, *_ = _ctx_for(gold)
    og = oracle.OracleGraph.load(gold["data_path"])
    for qi in (0, 1, 3, 6, 7):
        rec = gold["queries"][qi]
        if rec["limit"] is not None:
            continue
        qf = gold["query_paths_files"][qi]
        qo, qn, ql = gpe.host_load_graph(qf)
        cands = [np.array(c, dtype=np.uint32) for c in rec["candidates"]]
        res = ctx.refine(qo, qn, ql, cands, want_matches=rec["answer"] + 16)
        n, om = oracle.refine(og, oracle.OracleGraph.load(qf), cands, want_matches=rec["answer"] + 16)
        assert res["n_matches"] == n == rec["answer"]
        assert sorted(map(tuple, res["matches"].tolist())) == sorted(map(tuple, om.tolist()))
    ctx.close()


def test_sharded_tables_union_to_the_full_result():
    """Path table sharded by partition (what each GPU of a multi-GPU run holds): the union of the shards'
    candidate sets equals the unsharded result, and shard row counts are the reference's per-partition counts."""
    gold = load_case("quickstart")
    ctx, g, vde, sorted_nodes, membership, n_rows, rows_pp = _ctx_for(gold)
    L, e, p = gold["l"] + 1, gold["e"], gold["p"]
    rec = gold["queries"][0]
    qo, qn, ql = gpe.host_load_graph(gold["query_paths_files"][0])
    plan = gpe.host_query_plan(qo, qn, ql, L, e)
    union = [set() for _ in range(len(ql))]
    total = 0
    for shard in range(2):
        sel = np.array([1 if i % 2 == shard else 0 for i in range(p)], dtype=np.uint8)
        rows = ctx.build_table(sel)
        assert rows == sum(r for i, r in enumerate(gold["rows_per_partition"]) if i % 2 == shard)
        total += rows
        sets, _ = ctx.filter(plan, len(ql))
        for u, s in enumerate(sets):
            union[u] |= set(s.tolist())
    assert total == n_rows
    assert [sorted(s) for s in union] == rec["candidates"]
    ctx.close()


def test_edge_cases():
    gold = load_case("uniform300")
    ctx, g, *_ = _ctx_for(gold)
    ctx.build_table()
    # a 2-vertex query has no path of 3 vertices: empty plan, empty candidates, answer 0
    q2 = graph_io.csr_from_edges(2, np.array([[0, 1]]), np.array([0, 1]))
    assert ctx.query_batch([q2]).tolist() == [0]
    # a label that does not occur in the data graph
    q3 = graph_io.csr_from_edges(3, np.array([[0, 1], [1, 2]]), np.array([0, 99, 1]))
    assert ctx.query_batch([q3]).tolist() == [0]
    # disconnected queries are rejected (undefined behaviour in the reference)
    qd = graph_io.csr_from_edges(4, np.array([[0, 1], [2, 3]]), np.array([0, 1, 2, 3]))
    with pytest.raises(gpe.GpeError, match="disconnected"):
        ctx.query_batch([qd])
    # an edge listed from one end only
    qa = graph_io.csr_from_edges(3, np.array([[0, 1], [1, 2]]), np.array([0, 1, 2]))
    qa.nbrs[-1] = 0          # 2 lists 0, 0 does not list 2
    with pytest.raises(gpe.GpeError, match="symmetric"):
        ctx.query_batch([qa])
    # empty batch
    assert ctx.query_batch([]).tolist() == []
    # errors: wrong call order and unsupported shapes
    c2 = gpe.GpeContext(0)
    with pytest.raises(gpe.GpeError):
        c2.build_table()
    c2.set_graph(g.offsets, g.nbrs, g.labels)
    with pytest.raises(gpe.GpeError, match="unsupported"):
        c2.enumerate(6, graph_io.degree_order(g), graph_io.block_membership(g.V, 2), 2)
    bad = g.nbrs.copy()
    bad[0], bad[1] = bad[1], bad[0]
    with pytest.raises(gpe.GpeError, match="ascending"):
        c2.set_graph(g.offsets, bad, g.labels)
    c2.close()
    ctx.close()


@pytest.mark.parametrize("seed,V,E,nl,l,e", [(1, 2000, 12000, 5, 2, 2), (2, 1500, 6000, 3, 3, 4), (3, 3000, 30000, 8, 2, 8),
                                             (4, 5000, 20000, 200, 2, 1)])
def test_random_graphs_against_oracle(seed, V, E, nl, l, e):
    from oracle import oracle
    g = synth.chung_lu_graph(V, E, nl, gamma=2.6, degree_cap=80, seed=seed)
    sorted_nodes = graph_io.degree_order(g)
    p = 3
    membership = graph_io.block_membership(V, p)
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    n = og.enumerate(l + 1, sorted_nodes)
    ctx = gpe.GpeContext(0)
    ctx.set_graph(g.offsets, g.nbrs, g.labels)
    _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, e)
    assert vde.tobytes() == og.embeddings(e)[1].tobytes()
    ctx.set_embeddings(vde)
    n_rows, rows_pp = ctx.enumerate(l + 1, sorted_nodes, membership, p)
    assert n_rows == n
    assert rows_pp.tolist() == og.rows_per_partition(membership, p).tolist()
    if n < 3_000_000:
        assert np.array_equal(ctx.dump_paths(), og.paths())
    ctx.build_table()
    queries = synth.query_batch(g, 6, (4, 9), seed=seed + 100, mixed=True)
    expect = []
    for q in queries:
        oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
        sets, surv = oracle.filter_candidates(og, oq, e)
        plan = gpe.host_query_plan(q.offsets, q.nbrs, q.labels, l + 1, e)
        gsets, gsurv = ctx.filter(plan, q.V)
        assert [s.tolist() for s in gsets] == [s.tolist() for s in sets]
        assert gsurv.tolist() == surv.tolist()
        expect.append(oracle.refine(og, oq, sets, limit=2_000_000))
    ans = ctx.query_batch(queries, [2_000_000] * len(queries))
    assert ans.tolist() == expect
    ctx.close()


def test_host_cli_quickstart(tmp_path):
    """The C++ host keeps the reference's CLI, files and stdout: offline then online on the quick start."""
    import os
    import shutil
    import subprocess
    from tests.golden_util import ROOT
    main = os.path.join(ROOT, "host", "main")
    if not os.path.exists(main):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host")])
    gold = load_case("quickstart")
    d = str(tmp_path) + "/"
    os.makedirs(d + "gnn-pe/partitions")
    for i in range(gold["p"]):
        os.makedirs(d + f"gnn-pe/partitions/partition-{i}")
    shutil.copy(gold["membership_path"], d + "gnn-pe/membership.txt")
    out = subprocess.check_output([main, "-f", d, "-d", gold["data_path"], "-m", "offline"]).decode()
    assert "|V|: 3112, |E|: 12519, |Σ|: 71" in out and "Max Degree: 168, Max Label Frequency: 622" in out
    text = open(d + "gnn-pe/all_paths.txt", "rb").read()
    assert hashlib.md5(text).hexdigest() == gold["all_paths_md5"] == "61f074b1f90c96806c20f7000d1979a6"
    ids = []
    for i in range(gold["p"]):
        lines = open(d + f"gnn-pe/partitions/partition-{i}/partition_paths.txt").read().split()
        assert int(lines[0]) == gold["rows_per_partition"][i] == len(lines) - 1
        part = list(map(int, lines[1:]))
        assert part == sorted(part)
        ids += part
    assert sorted(ids) == list(range(gold["n_rows"]))
    out = subprocess.check_output([main, "--file", d, "--data", gold["data_path"], "--query",
                                   gold["query_paths_files"][0], "--mode=online", "-p5", "-l", "2", "-e", "2"]).decode()
    lines = out.strip().splitlines()
    assert lines[2] == "6"                       # plan size, custom.h:630
    assert lines[3].startswith("Answer Number: 45426 Query Time (ms): ")
    out = subprocess.check_output([main, "-f", d, "-d", gold["data_path"], "-q", gold["query_paths_files"][0],
                                   "-m", "online", "-n", "1000"]).decode()
    assert "Answer Number: 1000 " in out
    # a directory as -q: all its queries in one batch
    qdir = d + "queries"
    os.makedirs(qdir)
    for i in range(3):
        shutil.copy(gold["query_paths_files"][0], f"{qdir}/q{i}.graph")
    out = subprocess.check_output([main, "-f", d, "-d", gold["data_path"], "-q", qdir, "-m", "online"]).decode()
    assert out.count(": Answer Number: 45426") == 3 and "Queries: 3 Query Time (ms): " in out
    # the binary manifest of the offline run ties the outputs to (graph, membership.txt, -l, -p): other settings are refused
    assert os.path.getsize(d + "gnn-pe/paths.gpe") == 56 + 8 * gold["p"]
    r = subprocess.run([main, "-f", d, "-d", gold["data_path"], "-q", gold["query_paths_files"][0], "-m", "online", "-l", "3"],
                       capture_output=True)
    assert r.returncode == 1 and b"paths.gpe is stale (written for -l 2)" in r.stderr
    other = load_case("uniform300")
    r = subprocess.run([main, "-f", d, "-d", other["data_path"], "-q", gold["query_paths_files"][0], "-m", "online"],
                       capture_output=True)
    assert r.returncode == 1  # (membership.txt of another graph: rejected before or by the manifest)
    os.remove(d + "gnn-pe/paths.gpe")   # outputs of the reference's own offline run carry no manifest: nothing to check
    out = subprocess.check_output([main, "-f", d, "-d", gold["data_path"], "-q", gold["query_paths_files"][0], "-m", "online"]).decode()
    assert "Answer Number: 45426 " in out


def test_two_shards_in_one_process_match_single_gpu():
    """The multi-GPU data flow (shard tables -> export lists -> union -> split join -> sum) exercised with two
    contexts on one GPU, no NCCL: results must equal the unsharded run."""
    import torch
    gold = load_case("powerlaw500_e3")
    queries = [graph_io.read_graph(qf) for qf in gold["query_paths_files"]][:2] + \
              [graph_io.read_graph(qf) for qf in load_case("powerlaw500_e3")["query_paths_files"]][3:]
    want = [gold["queries"][i]["answer"] for i in (0, 1, 3)]
    limits = [gold["queries"][i]["limit"] or gpe.LIMIT_MAX for i in (0, 1, 3)]
    world = 2
    ctxs = []
    for r in range(world):
        ctx, g, vde, sorted_nodes, membership, n_rows, rows_pp = _ctx_for(gold)
        sel = np.array([1 if i % world == r else 0 for i in range(gold["p"])], dtype=np.uint8)
        ctx.build_table(sel)
        ctx.batch_upload(queries, limits)
        ctx.batch_filter()
        ctxs.append(ctx)
    infos = [c.batch_cand_info() for c in ctxs]
    n_slots = infos[0][0]
    stride = max(max(t for _, t in infos), 1)
    all_counts = torch.zeros(world, n_slots, dtype=torch.int32, device="cuda")
    all_cand = torch.zeros(world, stride, dtype=torch.int32, device="cuda")
    for r, c in enumerate(ctxs):
        c.batch_cand_export(all_counts[r].data_ptr(), all_cand[r].data_ptr())
    torch.cuda.synchronize()
    raw = np.zeros(len(queries), dtype=np.uint64)
    for r, c in enumerate(ctxs):
        c.batch_cand_merge(world, all_counts.data_ptr(), all_cand.data_ptr(), stride)
        off, cand = c.batch_get_candidates()
        slot = 0
        for qi in (0, 1, 3):
            for cset in gold["queries"][qi]["candidates"]:
                assert cand[int(off[slot]):int(off[slot + 1])].tolist() == cset
                slot += 1
        c.batch_join(r, world)
        raw += c.batch_download()
    got = [ctxs[0].clamp(int(x), l) for x, l in zip(raw, limits)]
    assert got == want
    for c in ctxs:
        c.close()


def test_two_shards_bitmap_exchange_matches_single_gpu():
    """The same flow in the bitmap form the multi-GPU engine uses (scan -> gather the shards' bitmaps -> OR fused
    into the compaction -> split join -> sum), two contexts on one GPU standing in for two ranks."""
    import torch
    gold = load_case("powerlaw500_e3")
    qfiles = gold["query_paths_files"]
    queries = [graph_io.read_graph(qfiles[i]) for i in (0, 1, 3)]
    want = [gold["queries"][i]["answer"] for i in (0, 1, 3)]
    limits = [gold["queries"][i]["limit"] or gpe.LIMIT_MAX for i in (0, 1, 3)]
    world = 2
    ctxs, parts = [], []
    for r in range(world):
        ctx, g, vde, sorted_nodes, membership, n_rows, rows_pp = _ctx_for(gold)
        ctx.build_table(sharding.partitions_of_rank(gold["p"], r, world))
        ctx.batch_upload(queries, limits)
        ctx.batch_scan()
        ctx.sync()
        ptr, nbytes = ctx.batch_bitmap()
        parts.append(torch.as_tensor(sharding._DevMem(ptr, nbytes), device="cuda").clone())
        ctxs.append(ctx)
    assert parts[0].numel() == parts[1].numel() > 0
    all_bm = torch.stack(parts).contiguous()
    union = sharding.union_bitmaps_reference(all_bm.cpu().numpy())
    torch.cuda.synchronize()
    raw = np.zeros(len(queries), dtype=np.uint64)
    for r, c in enumerate(ctxs):
        c.batch_bitmap_merge(world, all_bm.data_ptr())
        ptr, nbytes = c.batch_bitmap()
        c.sync()
        assert np.array_equal(torch.as_tensor(sharding._DevMem(ptr, nbytes), device="cuda").cpu().numpy(), union)
        off, cand = c.batch_get_candidates()
        slot = 0
        for qi in (0, 1, 3):
            for cset in gold["queries"][qi]["candidates"]:
                assert cand[int(off[slot]):int(off[slot + 1])].tolist() == cset
                slot += 1
        c.batch_join(r, world)
        raw += c.batch_download()
    assert [ctxs[0].clamp(int(x), l) for x, l in zip(raw, limits)] == want
    for c in ctxs:
        c.close()


# Query shapes that exercise every branch of the join's counted tail (leaves with distinct labels, several
# same-label leaves on one pivot, two same-label leaves on two pivots, leaves that share a label with walked
# vertices, leaves next to cycles) and the subtree hand-over between threads.
_TAIL_SHAPES = {
    "star4_same": ([(0, 1), (0, 2), (0, 3), (0, 4)], [0, 1, 1, 1, 1]),
    "star5_mixed": ([(0, 1), (0, 2), (0, 3), (0, 4), (0, 5)], [0, 1, 1, 2, 2, 0]),
    "two_hubs_pair": ([(0, 1), (0, 2), (1, 3)], [0, 1, 2, 2]),
    "two_hubs_four_leaves": ([(0, 1), (0, 2), (0, 3), (1, 4), (1, 5)], [0, 1, 2, 2, 2, 2]),
    "path_with_pendants": ([(0, 1), (1, 2), (2, 3), (1, 4), (2, 5), (3, 6)], [0, 1, 0, 1, 0, 1, 2]),
    "triangle_pendants": ([(0, 1), (1, 2), (0, 2), (0, 3), (1, 4), (2, 5)], [0, 1, 2, 1, 2, 0]),
    "square_pendant": ([(0, 1), (1, 2), (2, 3), (0, 3), (2, 4), (2, 5)], [0, 1, 0, 1, 1, 1]),
    "caterpillar": ([(0, 1), (1, 2), (2, 3), (0, 4), (1, 5), (2, 6), (3, 7)], [0, 1, 2, 0, 1, 2, 0, 1]),
    "leaf_label_equals_pivot_neighbour": ([(0, 1), (1, 2), (1, 3), (3, 4)], [0, 1, 2, 2, 0]),
    "edge_plus_leaf": ([(0, 1), (1, 2)], [0, 1, 0]),
    # core ends with a repeated label that carry pendant subtrees of query-unique labels: counted with weights
    # (one such end; a pair on two pivots; an end whose label also sits on a walked vertex next to / away from its pivot)
    "weighted_end": ([(0, 1), (1, 2), (2, 3), (3, 4), (0, 5), (5, 6)], [0, 1, 2, 3, 0, 4, 5]),
    "weighted_pair": ([(0, 1), (1, 2), (2, 3), (0, 4), (3, 5), (5, 6)], [0, 1, 2, 0, 3, 4, 5]),
    "weighted_end_label_on_walk": ([(0, 1), (1, 2), (2, 3), (1, 4), (4, 5), (3, 6)], [0, 1, 0, 2, 0, 3, 4]),
    "weighted_three": ([(0, 1), (1, 2), (2, 3), (3, 4), (0, 5), (4, 6), (2, 7)], [0, 1, 0, 2, 0, 3, 4, 5]),
}


@pytest.mark.parametrize("nl,seed", [(2, 11), (3, 12), (5, 13)])
def test_join_counted_tail_equals_reference_enumeration(nl, seed):
    from oracle import oracle
    g = synth.chung_lu_graph(300, 1500, nl, gamma=2.4, degree_cap=60, seed=seed)
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    ctx = gpe.GpeContext(0)
    ctx.set_graph(g.offsets, g.nbrs, g.labels)
    rng = np.random.default_rng(seed)
    for name, (edges, labels) in _TAIL_SHAPES.items():
        labels = np.array(labels) % nl
        q = graph_io.csr_from_edges(len(labels), np.array(edges), labels)
        oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
        # candidate sets as a caller may supply them: right-label vertices thinned at random, plus a few
        # vertices of the wrong label (the reference only ever iterates the start vertex's set, unchecked)
        cands = []
        for u in range(q.V):
            right = np.nonzero(g.labels == labels[u])[0]
            keep = right[rng.random(len(right)) < rng.uniform(0.3, 1.0)]
            wrong = rng.choice(g.V, size=3, replace=False)
            cands.append(np.unique(np.concatenate([keep, wrong])).astype(np.uint32))
        expect = oracle.refine(og, oq, cands)
        res = ctx.refine(q.offsets, q.nbrs, q.labels, cands)
        order, pivot = oracle.matching_order(og, oq, [len(c) for c in cands])
        assert res["order"].tolist() == order.tolist(), name
        assert res["n_matches"] == min(expect, gpe.LIMIT_MAX), (name, res["n_matches"], expect)
        if 0 < expect <= 200_000:
            n, om = oracle.refine(og, oq, cands, want_matches=expect + 8)
            rm = ctx.refine(q.offsets, q.nbrs, q.labels, cands, want_matches=expect + 8)
            assert rm["n_matches"] == expect, name
            assert sorted(map(tuple, rm["matches"].tolist())) == sorted(map(tuple, om.tolist())), name
    ctx.close()


def _random_query(rng, nq, extra_edges, nl):
    """A connected query: random tree plus a few extra edges, labels drawn from nl labels."""
    edges = set()
    for v in range(1, nq):
        edges.add((int(rng.integers(0, v)), v))
    for _ in range(extra_edges):
        a, b = sorted(rng.choice(nq, size=2, replace=False).tolist())
        edges.add((a, b))
    return np.array(sorted(edges)), rng.integers(0, nl, size=nq)


# The whole online stage (filter -> candidates -> join) on query shapes that exercise the join's factorisation:
# peeled pendant subtrees (labels unique in the query) tabulated per data vertex, a walk that starts from the core's
# label class when the start vertex itself is peeled, counted leaves with repeated labels, cycles.  The oracle
# enumerates every embedding the way the reference does.
@pytest.mark.parametrize("nl,seed", [(4, 21), (9, 22), (16, 23)])
def test_factorised_join_equals_reference_enumeration(nl, seed):
    from oracle import oracle
    g = synth.chung_lu_graph(1200, 7000, nl, gamma=2.4, degree_cap=60, seed=seed)
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    ctx = gpe.GpeContext(0)
    ctx.set_graph(g.offsets, g.nbrs, g.labels)
    _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, 2)
    ctx.set_embeddings(vde)
    sorted_nodes = graph_io.degree_order(g)
    assert ctx.enumerate(3, sorted_nodes, graph_io.block_membership(g.V, 3), 3)[0] == og.enumerate(3, sorted_nodes)
    ctx.build_table()
    rng = np.random.default_rng(seed)
    queries, names = [], []
    for name, (edges, labels) in _TAIL_SHAPES.items():
        if len(labels) < 3:
            continue
        for variant in range(2):  # the shape's own label pattern, and one with all labels distinct where possible
            lab = np.array(labels) % nl if variant == 0 else rng.permutation(max(nl, len(labels)))[:len(labels)] % nl
            queries.append(graph_io.csr_from_edges(len(lab), np.array(edges), lab))
            names.append(f"{name}/{variant}")
    for i in range(24):
        nq = int(rng.integers(3, 10))
        edges, lab = _random_query(rng, nq, int(rng.integers(0, 3)), nl)
        queries.append(graph_io.csr_from_edges(nq, edges, lab))
        names.append(f"random{i}")
    limit = 50_000_000
    expect = [oracle.online(og, oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels), 2, limit) for q in queries]
    ans = ctx.query_batch(queries, [limit] * len(queries)).tolist()
    bad = [(n, a, e) for n, a, e in zip(names, ans, expect) if a != e]
    assert not bad, bad
    assert sum(1 for e in expect if e > 0) >= len(expect) // 4  # the cases are not vacuous
    # one query at a time gives the same answers (different batch composition, same plans)
    for q, e in list(zip(queries, expect))[::5]:
        assert int(ctx.query_batch([q], [limit])[0]) == e
    ctx.close()


# Counted leaves that carry peeled subtrees (weighted sums and pairs, the default) against the same leaves walked
# (GPE_JOIN_WEIGHTED=0, also what a query falls back to when a weighted count saturates) and against the oracle's
# enumeration, on one batch of query shapes.
@pytest.mark.parametrize("nl,seed", [(9, 31), (16, 32)])
def test_weighted_leaves_equal_walked_leaves_and_oracle(nl, seed, monkeypatch):
    from oracle import oracle
    g = synth.chung_lu_graph(1200, 7000, nl, gamma=2.4, degree_cap=60, seed=seed)
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    sorted_nodes = graph_io.degree_order(g)
    og.enumerate(3, sorted_nodes)
    rng = np.random.default_rng(seed)
    queries = []
    for name, (edges, labels) in _TAIL_SHAPES.items():
        if len(labels) >= 3:
            queries.append(graph_io.csr_from_edges(len(labels), np.array(edges), np.array(labels) % nl))
    for i in range(24):
        nq = int(rng.integers(3, 10))
        edges, lab = _random_query(rng, nq, int(rng.integers(0, 3)), nl)
        queries.append(graph_io.csr_from_edges(nq, edges, lab))
    cap = 20_000_000  # bounds the oracle's enumeration; queries that reach it are left out (no limit may be set here)
    expect = [oracle.online(og, oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels), 2, cap) for q in queries]
    queries = [q for q, e in zip(queries, expect) if e < cap]
    expect = [e for e in expect if e < cap]
    assert len(queries) >= 20 and sum(1 for e in expect if e > 0) >= len(expect) // 4

    def run(env):
        for k in ("GPE_JOIN_WEIGHTED", "GPE_JOIN_POOL", "GPE_JOIN_PEEL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ctx = gpe.GpeContext(0)
        ctx.set_graph(g.offsets, g.nbrs, g.labels)
        _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, 2)
        ctx.set_embeddings(vde)
        ctx.enumerate(3, sorted_nodes, graph_io.block_membership(g.V, 3), 3)
        ctx.build_table()
        ans = ctx.query_batch(queries).tolist()
        st = ctx.stats()
        one = [int(ctx.query_batch([q])[0]) for q in queries[::7]]
        ctx.close()
        return ans, st, one

    weighted, st_w, one_w = run({})
    bad = [(i, a, e) for i, (a, e) in enumerate(zip(weighted, expect)) if a != e]
    assert not bad, bad
    assert one_w == expect[::7]
    walked, st_d, _ = run({"GPE_JOIN_WEIGHTED": "0"})
    assert walked == expect
    assert st_d["join_steps"] >= st_w["join_steps"]  # counting never walks more than walking
    # a table pool too small for any table (a query whose tables do not fit walks instead) and no tables at all
    tiny, st_t, _ = run({"GPE_JOIN_POOL": "64"})
    assert tiny == expect and st_t["join_steps"] >= st_w["join_steps"]
    none, _, _ = run({"GPE_JOIN_PEEL": "0"})
    assert none == expect
