"""Join utilisation on a bench workload: the whole batch, then every query alone (time, steps, lane utilisation).
   usage: python tools/joinstats.py <workload> [top]"""
import os, sys, time, numpy as np
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from gnn_pe_b200 import gpe, graph_io

name = sys.argv[1] if len(sys.argv) > 1 else "small"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 15
w, g, queries = bench.load_workload(name)
ctx = gpe.GpeContext(0)
ctx.set_graph(g.offsets, g.nbrs, g.labels)
_, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, w["e"])
ctx.set_embeddings(vde)
ctx.enumerate(w["l"] + 1, graph_io.degree_order(g), graph_io.block_membership(g.V, w["p"]), w["p"])
ctx.build_table()
ctx.set_timing(1)


def line(tag, st, a):
    wi, st_ = max(st["join_warp_iters"], 1), st["join_steps"]
    return (f"{tag}: join={st['last_join_ms']:.3f} ms scan={st['last_scan_ms']:.3f} ms matches={a} steps={st_} "
            f"warp_iters={wi} lane_util={st_ / (32 * wi):.3f} idle_polls={st['join_idle_polls']} "
            f"items={st['join_items']} exports={st['join_exports']} donations={st['join_donations']} cand={st['n_candidates']}")


for _ in range(3):
    a = ctx.query_batch(queries)
    st = ctx.stats()
    print(line("batch", st, int(a.sum())))
rows = []
for i, q in enumerate(queries if top > 0 else []):
    ctx.query_batch([q])
    a = int(ctx.query_batch([q])[0])
    st = ctx.stats()
    rows.append((st["last_join_ms"], i, line(f"q{i:03d} nq={q.V} ne={len(q.nbrs)//2}", st, a)))
print("sum of single-query join ms:", sum(r[0] for r in rows))
for r in sorted(rows, reverse=True)[:top]:
    print(r[2])
ctx.close()
