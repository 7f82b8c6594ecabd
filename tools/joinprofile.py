"""Per-query join statistics (answers, DFS steps, trailing-leaf shortcut) for a bench workload."""
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
ctx.enumerate(w["l"] + 1, graph_io.degree_order(g), graph_io.block_membership(g.V, 8), 8)
ctx.build_table()
rows = []
for i, q in enumerate(queries):
    a = int(ctx.query_batch([q])[0])
    st = ctx.stats()
    order, pivot = ctx.batch_get_plan(q.V)
    deg = q.degrees
    depth_of = {int(u): d for d, u in enumerate(order)}
    nq = q.V
    fast, pvd, lab = [], [], []
    for d, u in enumerate(order):
        u = int(u)
        nb = [int(x) for x in q.nbrs[q.offsets[u]:q.offsets[u + 1]]]
        pv = int(pivot[d]) if d else -1
        bn = [x for x in nb if depth_of[x] < d and x != pv]
        fast.append(d > 0 and not bn and deg[u] <= 1)
        pvd.append(depth_of[pv] if d else 0)
        lab.append(int(q.labels[u]))
    k = 0
    while k + 1 < nq:
        t = nq - 1 - k
        ok = fast[t] and all(pvd[j] < t for j in range(t, nq)) and all(lab[j] != lab[t] for j in range(t + 1, nq))
        if not ok:
            break
        k += 1
    off, cand = ctx.batch_get_candidates()
    cnt = np.diff(off.astype(np.int64))
    rows.append((a, st["join_steps"], st["join_exports"], k, int(cnt[int(order[0])]), [int(deg[int(u)]) for u in order]))
rows_sorted = sorted(rows, key=lambda r: -r[1])
tot_steps = sum(r[1] for r in rows)
print("total matches", sum(r[0] for r in rows), "total steps", tot_steps)
print("tail_k histogram", np.bincount([r[3] for r in rows]))
for r in rows_sorted[:15]:
    print(f"matches={r[0]:>12} steps={r[1]:>11} ({100*r[1]/tot_steps:4.1f}%) rounds={r[2]} tail_k={r[3]} |C(start)|={r[4]} qdeg_in_order={r[5]}")
