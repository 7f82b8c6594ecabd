#!/usr/bin/env python
"""Generate the committed golden vectors by running the UNMODIFIED reference here.

Needs oracle/_ref/main and oracle/_ref/probe (``make -C oracle``, which needs /root/reference).
Run from the repo root:  python tests/golden/make_golden.py
Outputs, per case, tests/golden/<case>/{data.graph, membership.txt, q*.graph, golden.json}.

What produced each value:
  * ``all_paths_md5`` / ``n_rows`` / ``rows_per_partition``: ``main -m offline`` (unmodified binary) for
    l=2; for l!=2 the probe's ``offline`` mode, i.e. the reference's ``dfs`` started at depth 1
    (SURVEY.md F5 -- the unmodified binary has no defined behaviour there).
  * ``main_answer``: the ``Answer Number`` the unmodified ``main -m online`` prints (l=2 cases).
  * everything else: oracle/_ref/probe, which calls the reference's own gen_vde / gen_pde /
    Partition (R*-tree build + traversal) / gen_query_pde / generateGQLQueryPlan / refinement.
"""
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gnn_pe_b200 import graph_io, synth  # noqa: E402

REF_MAIN = os.path.join(ROOT, "oracle/_ref/main")
REF_PROBE = os.path.join(ROOT, "oracle/_ref/probe")
OUT = os.path.join(ROOT, "tests/golden")


def cycle_query(labels):
    n = len(labels)
    edges = [(i, (i + 1) % n) for i in range(n)]
    return graph_io.csr_from_edges(n, np.array(edges), np.array(labels))


def run_case(name, g, queries, p, l, e, limits=None):
    case_dir = os.path.join(OUT, name)
    os.makedirs(case_dir, exist_ok=True)
    work = tempfile.mkdtemp(prefix="golden_") + "/"
    data_path = os.path.join(case_dir, "data.graph")
    if not os.path.exists(data_path):
        graph_io.write_graph(data_path, g)
    sorted_nodes = graph_io.degree_order(g)
    membership = graph_io.block_membership(g.V, p)
    os.makedirs(work + "gnn-pe/partitions", exist_ok=True)
    for i in range(p):
        os.makedirs(work + f"gnn-pe/partitions/partition-{i}", exist_ok=True)
    graph_io.write_membership(work + "gnn-pe/membership.txt", sorted_nodes, membership)
    shutil.copy(work + "gnn-pe/membership.txt", os.path.join(case_dir, "membership.txt"))

    unmodified = (l == 2)
    if unmodified:
        subprocess.check_call([REF_MAIN, "-f", work, "-d", data_path, "-m", "offline", "-p", str(p),
                               "-l", str(l), "-e", str(e)], stdout=subprocess.DEVNULL)
    else:
        subprocess.check_call([REF_PROBE, "offline", work, data_path, str(p), str(l), "1"])
    text = open(work + "gnn-pe/all_paths.txt", "rb").read()
    n_rows = int(text.split(b"\n", 1)[0])
    rows_per_partition = [int(open(work + f"gnn-pe/partitions/partition-{i}/partition_paths.txt").readline())
                          for i in range(p)]
    first_rows = [list(map(int, ln.split())) for ln in text.split(b"\n")[1:4] if ln.strip()]
    gold = dict(name=name, p=p, l=l, e=e, unmodified_reference=unmodified, n_rows=n_rows,
                all_paths_md5=hashlib.md5(text).hexdigest(), rows_per_partition=rows_per_partition,
                first_rows=first_rows, queries=[])
    for qi, q in enumerate(queries):
        qpath = os.path.join(case_dir, f"q{qi}.graph")
        graph_io.write_graph(qpath, q)
        limit = (limits or {}).get(qi)
        args = [REF_PROBE, work, data_path, qpath, str(p), str(l), str(e), "-1" if unmodified else "1"]
        if limit is not None:
            args.append(str(limit))
        out = subprocess.check_output(args).decode()
        rec = json.loads(out.split("=====JSON=====", 1)[1])
        rec["limit"] = limit
        if unmodified:
            margs = [REF_MAIN, "-f", work, "-d", data_path, "-q", qpath, "-m", "online", "-p", str(p),
                     "-l", str(l), "-e", str(e)]
            if limit is not None:
                margs += ["-n", str(limit)]
            mout = subprocess.check_output(margs).decode()
            rec["main_answer"] = int(re.search(r"Answer Number: (\d+)", mout).group(1))
            assert rec["main_answer"] == rec["answer"], (name, qi, rec["main_answer"], rec["answer"])
        assert rec["index_equals_brute"], (name, qi, "index traversal != all-pairs compare")
        gold["queries"].append(rec)
        print(f"  {name} q{qi}: nq={q.V} plan={rec['plan_size']} |C|={rec['candidate_counts']} answer={rec['answer']}")
    with open(os.path.join(case_dir, "golden.json"), "w") as f:
        json.dump(gold, f, indent=0, separators=(",", ":"))
    shutil.rmtree(work)
    print(f"{name}: rows={n_rows} md5={gold['all_paths_md5']}")


def main():
    # 1. the reference's own quick start (BASELINE.json configs[0]); inputs are the shipped Test/ files
    qs_dir = os.path.join(OUT, "quickstart")
    os.makedirs(qs_dir, exist_ok=True)
    if os.path.exists("/root/reference/Test/data_graph.graph"):
        shutil.copy("/root/reference/Test/data_graph.graph", os.path.join(qs_dir, "data.graph"))
        shutil.copy("/root/reference/Test/query_graph.graph", os.path.join(qs_dir, "q0_shipped.graph"))
    g = graph_io.read_graph(os.path.join(qs_dir, "data.graph"))
    q0 = graph_io.read_graph(os.path.join(qs_dir, "q0_shipped.graph"))
    rng = np.random.default_rng(7)
    queries = [q0] + [synth.random_walk_query(g, n, rng, induced=ind) for n, ind in
                      [(5, True), (6, True), (8, False), (10, True), (4, True)]]
    run_case("quickstart", g, queries, p=5, l=2, e=2, limits={3: 100})

    # 2. small uniform graph, few labels (dense candidate sets), cyclic + 12-vertex queries
    g = synth.uniform_graph(300, 1200, 4, seed=11)
    rng = np.random.default_rng(12)
    queries = [synth.random_walk_query(g, n, rng, induced=ind) for n, ind in
               [(5, True), (6, True), (8, False), (12, True), (12, False), (3, True)]]
    queries.append(cycle_query([0, 1, 2, 3]))
    queries.append(cycle_query([1, 1, 1]))
    run_case("uniform300", g, queries, p=3, l=2, e=2, limits={1: 7, 2: 1})

    # 3. power-law graph, e=3 (odd embedding width), p=4
    g = synth.chung_lu_graph(500, 2500, 3, gamma=2.5, degree_cap=60, seed=21)
    rng = np.random.default_rng(22)
    queries = [synth.random_walk_query(g, n, rng, induced=ind) for n, ind in
               [(6, True), (7, True), (9, False), (16, True)]]
    run_case("powerlaw500_e3", g, queries, p=4, l=2, e=3, limits={1: 1000})

    # 4. l=3, e=4 (BASELINE.json configs[3] shape) -- patched-oracle semantics, SURVEY.md F5
    g = synth.uniform_graph(200, 600, 3, seed=31)
    rng = np.random.default_rng(32)
    queries = [synth.random_walk_query(g, n, rng, induced=ind) for n, ind in
               [(5, True), (6, False), (8, True), (12, True)]]
    run_case("uniform200_l3e4", g, queries, p=2, l=3, e=4)


if __name__ == "__main__":
    main()
