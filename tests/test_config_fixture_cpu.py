"""tests/golden/config_answers.json: the fixture bench.py and the GPU tests compare sampled queries with.  Here (CPU):
it covers every workload bench.py names, its small cases are reproduced by the oracle, and the generators still
produce the graphs it was made for."""
import json
import os

import bench
from gnn_pe_b200 import graph_io
from oracle import oracle
from tests.golden_util import ROOT

FIX = json.load(open(os.path.join(ROOT, "tests", "golden", "config_answers.json")))


def test_fixture_covers_the_bench_workloads():
    for name in ("config2", "config3", "config5", "small", "config3_small", "config5_small"):
        assert name in FIX and len(FIX[name]["answers"]) >= 5, name
        assert FIX[name]["desc"] == bench.WORKLOADS[name]["desc"]


def test_small_workloads_reproduce():
    for name in ("config3_small", "config5_small"):
        w, g, queries = bench.load_workload(name)
        rec = FIX[name]
        assert (g.V, g.E, len(queries)) == (rec["V"], rec["E"], rec["n_queries"])
        og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
        sorted_nodes = graph_io.degree_order(g)
        _, vde = og.embeddings(w["e"])
        for i, want in list(rec["answers"].items())[:4]:
            q = queries[int(i)]
            assert [q.V, q.E] == rec["query_sizes"][i]
            oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
            n, _ = oracle.online_streaming(og, oq, w["l"] + 1, w["e"], sorted_nodes, vde, threads=4)
            assert n == want
