"""Survivor statistics of the dominance filter on a bench workload: per query, (plan path, data path) pairs that pass
the leaf compare vs. distinct candidates produced.  usage: python tools/survivors.py <workload>"""
import os, sys, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from gnn_pe_b200 import gpe, graph_io

name = sys.argv[1] if len(sys.argv) > 1 else "small"
w, g, queries = bench.load_workload(name)
ctx = gpe.GpeContext(0)
ctx.set_graph(g.offsets, g.nbrs, g.labels)
_, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, w["e"])
ctx.set_embeddings(vde)
ctx.enumerate(w["l"] + 1, graph_io.degree_order(g), graph_io.block_membership(g.V, w["p"]), w["p"])
ctx.build_table()
tot_s = tot_c = tot_rows = 0
for q in queries:
    plan = gpe.host_query_plan(q.offsets, q.nbrs, q.labels, w["l"] + 1, w["e"])
    sets, surv = ctx.filter(plan, q.V)
    st = ctx.stats()
    tot_s += int(surv.sum()); tot_c += sum(len(s) for s in sets); tot_rows += st["scan_rows"]
print(f"rows examined {tot_rows}  survivor pairs {tot_s} ({tot_s / max(tot_rows, 1):.3f} per row)  distinct candidates {tot_c}  "
      f"bit sets requested {tot_s * (w['l'] + 1)}")
ctx.close()
