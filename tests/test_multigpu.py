"""N = 2 GPUs (run with `gpurun --gpus 2`; skipped on a 1-GPU box): the library's NCCL path in both of its forms --
one process per GPU (bench.py under torchrun) and one process driving two contexts (gpe_comm_init_all /
gpe_multi_query_batch, what `host/main -g 2` does) -- against the 1-GPU answers and the golden vectors."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.golden_util import ROOT, load_case

pytestmark = pytest.mark.gpu


def _two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def test_bench_two_gpus_matches_one_gpu():
    _two_gpus()
    env = dict(os.environ, GPE_BENCH_WORKLOAD="small")
    one = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3",
                                   "--no-cpu-baseline"], env=env).decode().strip().splitlines()[-1]
    two = subprocess.check_output([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "bench.py"),
                                   "--gpus", "2", "--steps", "2", "--warmup", "3", "--no-cpu-baseline"],
                                  env=env).decode().strip().splitlines()[-1]
    a, b = json.loads(one), json.loads(two)
    assert a["answers_checksum"] == b["answers_checksum"] and a["answers_nonzero"] == b["answers_nonzero"]
    assert b["n_gpus"] == 2 and b["nccl"]["ranks"] == 2
    assert a["oracle_parity"]["ok"] and b["oracle_parity"]["ok"] and b["oracle_parity"]["checked"] >= 5


@pytest.mark.parametrize("name,exchange,cap", [("quickstart", "dense", None), ("uniform300", "dense", None),
                                               ("quickstart", "sparse", None), ("uniform300", "sparse", "8")])
def test_one_process_two_contexts_gives_the_golden_answers(name, exchange, cap, monkeypatch):
    _two_gpus()
    monkeypatch.setenv("GPE_EXCHANGE", exchange)  # both forms of the candidate exchange; cap 8: sparse buffer too small, redone dense
    if cap:
        monkeypatch.setenv("GPE_SPARSE_CAP", cap)
    from gnn_pe_b200 import gpe, graph_io
    gold = load_case(name)
    g = graph_io.read_graph(gold["data_path"])
    sorted_nodes, membership = graph_io.read_membership(gold["membership_path"], g.V)
    _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, gold["e"])
    multi = gpe.MultiGpu([0, 1])
    try:
        n_rows, shard_rows = multi.build(g, gold["l"] + 1, gold["e"], gold["p"], sorted_nodes, membership, vde)
        assert n_rows == gold["n_rows"] and sum(shard_rows) == n_rows
        assert shard_rows[0] == sum(gold["rows_per_partition"][0::2]) and shard_rows[1] == sum(gold["rows_per_partition"][1::2])
        queries = [graph_io.read_graph(qf) for qf in gold["query_paths_files"]]
        limits = [r["limit"] if r["limit"] is not None else gpe.LIMIT_MAX for r in gold["queries"]]
        for _ in range(2):
            assert multi.query_batch(queries, limits).tolist() == [r["answer"] for r in gold["queries"]]
        # every GPU ends with the same (full) candidate sets: the exchange is the reference's merge (main.cpp:166-172)
        for c in multi.ctxs:
            off, cand = c.batch_get_candidates()
            slot = 0
            for r in gold["queries"]:
                for cset in r["candidates"]:
                    assert cand[int(off[slot]):int(off[slot + 1])].tolist() == cset
                    slot += 1
    finally:
        multi.close()


def test_host_cli_two_gpus(tmp_path):
    _two_gpus()
    import shutil
    gold = load_case("quickstart")
    d = tmp_path / "ds"
    (d / "gnn-pe" / "partitions").mkdir(parents=True)
    for i in range(gold["p"]):
        (d / "gnn-pe" / "partitions" / f"partition-{i}").mkdir()
    shutil.copy(gold["membership_path"], d / "gnn-pe" / "membership.txt")
    exe = os.path.join(ROOT, "host", "main")
    common = [exe, "-f", str(d) + "/", "-d", gold["data_path"], "-q", gold["query_paths_files"][0], "-p", str(gold["p"])]
    one = subprocess.check_output(common + ["-m", "online"]).decode()
    two = subprocess.check_output(common + ["-m", "online", "-g", "2"]).decode()
    assert "Answer Number: 45426" in one and "Answer Number: 45426" in two
