"""Short, deterministic target for ncu: builds a bench workload, then runs
   [steps] x (filter + join) over the query batch and [stream] streaming (no-prune) scans.
   usage: python tools/profile_target.py <workload> [steps] [stream_reps]"""
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from gnn_pe_b200 import gpe, graph_io

name = sys.argv[1] if len(sys.argv) > 1 else "small"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
stream_reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
w, g, queries = bench.load_workload(name)
ctx = gpe.GpeContext(0)
ctx.set_graph(g.offsets, g.nbrs, g.labels)
_, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, w["e"])
ctx.set_embeddings(vde)
ctx.enumerate(w["l"] + 1, graph_io.degree_order(g), graph_io.block_membership(g.V, w["p"]), w["p"])
ctx.build_table()
ctx.batch_upload(queries)
ctx.batch_filter()
ctx.batch_join()
ctx.batch_download()  # settles the candidate buffer (sized before the first total is known): later steps are the steady state
for _ in range(steps):
    ctx.batch_filter()
    ctx.batch_join()
ans = ctx.batch_download()
st = ctx.stats()
print("answers checksum", int(ans.sum()), {k: st[k] for k in ("scan_items", "scan_rows", "n_candidates", "join_items", "join_exports", "join_steps", "kernel_launches")})
q0 = queries[0]
plan = gpe.host_query_plan(q0.offsets, q0.nbrs, q0.labels, w["l"] + 1, w["e"])
for _ in range(stream_reps):
    ctx.filter(plan, q0.V, gpe.FILTER_NO_PRUNE)
print("streaming rows", ctx.stats()["scan_rows"])
ctx.close()
