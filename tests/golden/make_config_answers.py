"""Answers of sampled queries of bench.py's full-size workloads, computed by the oracle's streaming restatement
(oracle.online_streaming: all-pairs leaf compare walked from the CSR + the reference's refinement) on the CPU.
The GPU box has no time budget for this (a config-3 query takes ~1 min of host time), so the expected answers travel
as a fixture: tests/golden/config_answers.json.  bench.py and tests/test_gpu_configs.py compare against it.

    python tests/golden/make_config_answers.py config2 config3 config5     # ~15 min on 8 cores
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from gnn_pe_b200 import graph_io  # noqa: E402
from oracle import oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "config_answers.json")
SAMPLES = {"config2": [0, 17, 38, 59, 80, 99], "config3": [0, 21, 42, 63, 84], "config5": [0, 201, 402, 603, 804, 999],
           "config4": [0, 33, 66], "config4_small": list(range(0, 50, 7)),
           "config3_small": list(range(0, 100, 9)), "config5_small": list(range(0, 200, 13)), "small": list(range(0, 100, 11))}


def main():
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in sys.argv[1:]:
        w, g, queries = bench.load_workload(name)
        og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
        sorted_nodes = graph_io.degree_order(g)
        _, vde = og.embeddings(w["e"])
        ans = {}
        for i in SAMPLES[name]:
            q = queries[i]
            oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
            t0 = time.time()
            n, _ = oracle.online_streaming(og, oq, w["l"] + 1, w["e"], sorted_nodes, vde, threads=os.cpu_count())
            ans[str(i)] = int(n)
            print(name, i, q.V, q.E, n, f"{time.time() - t0:.1f}s", flush=True)
        res[name] = dict(desc=w["desc"], V=g.V, E=g.E, n_queries=len(queries), limit="UINT_MAX", answers=ans,
                         query_sizes={str(i): [int(queries[i].V), int(queries[i].E)] for i in SAMPLES[name]})
        json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
