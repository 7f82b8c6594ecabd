import sys, os, time, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import bench
from gnn_pe_b200 import gpe, graph_io
name = sys.argv[1] if len(sys.argv) > 1 else "small"
w, g, queries = bench.load_workload(name)
ctx = gpe.GpeContext(0)
ctx.set_graph(g.offsets, g.nbrs, g.labels)
_, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, 2)
ctx.set_embeddings(vde)
ctx.enumerate(3, graph_io.degree_order(g), graph_io.block_membership(g.V, 8), 8)
ctx.build_table()
for i in range(2):
    t=time.time(); a = ctx.query_batch(queries); dt=time.time()-t
    st = ctx.stats()
    print(f"batch: {dt*1e3:.2f} ms  matches={int(a.sum())} rounds={st['join_exports']} steps={st['join_steps']} items={st['join_items']}", file=sys.stderr)
