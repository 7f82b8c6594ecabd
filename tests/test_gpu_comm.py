"""The library's own NCCL path (gpe_comm_*, gpe_batch_step / gpe_batch_finish) on ONE GPU: a communicator of a single
rank takes the same route as N ranks -- scan, ncclAllGather of the candidate bitmaps, fused merge + compaction, join,
ncclAllReduce of the counts -- so the exchange code runs in the 1-GPU parity suite.  N > 1: tests/test_multigpu.py."""
import numpy as np
import pytest

from gnn_pe_b200 import gpe, graph_io
from tests.golden_util import CASES, load_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASES)
def test_single_rank_communicator_gives_the_golden_answers(name):
    gold = load_case(name)
    g = graph_io.read_graph(gold["data_path"])
    sorted_nodes, membership = graph_io.read_membership(gold["membership_path"], g.V)
    ctx = gpe.GpeContext(0)
    try:
        ctx.comm_init(0, 1, gpe.comm_unique_id())
        rank, world, version = ctx.comm_info()
        assert (rank, world) == (0, 1) and version >= 21800  # NCCL >= 2.18 loaded at run time
        ctx.set_graph(g.offsets, g.nbrs, g.labels)
        _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, gold["e"])
        ctx.set_embeddings(vde)
        n_rows, _ = ctx.enumerate(gold["l"] + 1, sorted_nodes, membership, gold["p"])
        assert ctx.build_table_shard() == n_rows == gold["n_rows"]
        queries = [graph_io.read_graph(qf) for qf in gold["query_paths_files"]]
        limits = [r["limit"] if r["limit"] is not None else gpe.LIMIT_MAX for r in gold["queries"]]
        ctx.batch_upload(queries, limits)
        for _ in range(2):  # a step can be repeated on an uploaded batch
            ctx.batch_step()
        assert ctx.batch_finish().tolist() == [r["answer"] for r in gold["queries"]]
        off, cand = ctx.batch_get_candidates()
        slot = 0
        for r in gold["queries"]:
            for cset in r["candidates"]:
                assert cand[int(off[slot]):int(off[slot + 1])].tolist() == cset
                slot += 1
    finally:
        ctx.close()


def test_candidate_buffer_overflow_is_repaired():
    """The candidate lists are expanded into a buffer sized before their total is known (no host sync in the step);
    a batch that outgrows it is redone transparently.  A tiny first batch sets a tiny capacity; the next batch is large."""
    gold = load_case("quickstart")
    g = graph_io.read_graph(gold["data_path"])
    sorted_nodes, membership = graph_io.read_membership(gold["membership_path"], g.V)
    ctx = gpe.GpeContext(0)
    try:
        ctx.set_graph(g.offsets, g.nbrs, g.labels)
        _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, 2)
        ctx.set_embeddings(vde)
        ctx.enumerate(3, sorted_nodes, membership, gold["p"])
        ctx.build_table()
        q = graph_io.read_graph(gold["query_paths_files"][0])
        import os
        os.environ["GPE_CAND_CAP"] = "16"  # entries: far below the 675 candidates of this query
        try:
            assert int(ctx.query_batch([q])[0]) == 45426
            off, cand = ctx.batch_get_candidates()
            assert [cand[int(off[i]):int(off[i + 1])].tolist() for i in range(q.V)] == gold["queries"][0]["candidates"]
            ctx.batch_upload([q, q])
            ctx.batch_filter()
            off2, cand2 = ctx.batch_get_candidates()  # settled by the getter, before any join
            assert int(off2[-1]) == 2 * int(off[-1])
        finally:
            del os.environ["GPE_CAND_CAP"]
    finally:
        ctx.close()


@pytest.mark.parametrize("mode,cap", [("sparse", None), ("sparse", "8"), ("dense", None)])
def test_exchange_forms_agree(mode, cap, monkeypatch):
    """Dense bitmaps, sparse (index, word) pairs, and a sparse buffer that is too small (the step is then redone with the
    dense exchange): same answers.  One rank here; tests/test_multigpu.py runs two."""
    gold = load_case("uniform300")
    g = graph_io.read_graph(gold["data_path"])
    sorted_nodes, membership = graph_io.read_membership(gold["membership_path"], g.V)
    monkeypatch.setenv("GPE_EXCHANGE", mode)
    if cap:
        monkeypatch.setenv("GPE_SPARSE_CAP", cap)
    ctx = gpe.GpeContext(0)
    try:
        ctx.comm_init(0, 1, gpe.comm_unique_id())
        ctx.set_graph(g.offsets, g.nbrs, g.labels)
        _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, gold["e"])
        ctx.set_embeddings(vde)
        ctx.enumerate(gold["l"] + 1, sorted_nodes, membership, gold["p"])
        ctx.build_table_shard()
        queries = [graph_io.read_graph(qf) for qf in gold["query_paths_files"]]
        limits = [r["limit"] if r["limit"] is not None else gpe.LIMIT_MAX for r in gold["queries"]]
        ctx.batch_upload(queries, limits)
        ctx.batch_step()
        assert ctx.batch_finish().tolist() == [r["answer"] for r in gold["queries"]]
        assert ctx.stats()["exchange_redos"] == (1 if cap else 0)
    finally:
        ctx.close()


@pytest.mark.parametrize("with_comm", [False, True])
def test_pipelined_batches_equal_single_calls(with_comm):
    """gpe_query_batches (host planning of batch i+1 behind the GPU work of batch i) returns what gpe_query_batch returns
    batch by batch -- batches of different composition, with limits, with and without a (one-rank) communicator."""
    gold = load_case("uniform300")
    g = graph_io.read_graph(gold["data_path"])
    sorted_nodes, membership = graph_io.read_membership(gold["membership_path"], g.V)
    ctx = gpe.GpeContext(0)
    try:
        if with_comm:
            ctx.comm_init(0, 1, gpe.comm_unique_id())
        ctx.set_graph(g.offsets, g.nbrs, g.labels)
        _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, gold["e"])
        ctx.set_embeddings(vde)
        ctx.enumerate(gold["l"] + 1, sorted_nodes, membership, gold["p"])
        ctx.build_table_shard()
        queries = [graph_io.read_graph(qf) for qf in gold["query_paths_files"]]
        limits = [r["limit"] if r["limit"] is not None else gpe.LIMIT_MAX for r in gold["queries"]]
        want = [r["answer"] for r in gold["queries"]]
        batches = [queries, queries[:3], queries[::-1], queries[2:]]
        lims = [limits, limits[:3], limits[::-1], limits[2:]]
        got = ctx.query_batches(batches, lims)
        assert [a.tolist() for a in got] == [want, want[:3], want[::-1], want[2:]]
        assert ctx.query_batches([]) == []
    finally:
        ctx.close()
