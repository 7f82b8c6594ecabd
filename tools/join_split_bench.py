"""The join of ONE rank of an N-way split, timed on one GPU: after the exchange every GPU holds the same candidate
sets, so `gpe_batch_join(rank, world)` on a single GPU with the full table does exactly the work rank `rank` of `world`
GPUs does.  Lets the scaling behaviour of the join (tail, hand-over cadence, replicated setup) be A/B-ed without
multi-GPU minutes.   usage: python tools/join_split_bench.py [workload] [reps]     (environment switches apply)"""
import json, os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
from gnn_pe_b200 import gpe, graph_io

name = sys.argv[1] if len(sys.argv) > 1 else "config2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
w, g, queries = bench.load_workload(name)
ctx = gpe.GpeContext(0)
ctx.set_graph(g.offsets, g.nbrs, g.labels)
_, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, w["e"])
ctx.set_embeddings(vde)
ctx.enumerate(w["l"] + 1, graph_io.degree_order(g), graph_io.block_membership(g.V, w["p"]), w["p"])
ctx.build_table()
ctx.batch_upload(queries)
ctx.batch_filter()
ctx.batch_join(0, 1)
total = ctx.batch_download()
out = {}
for world in (1, 2, 4, 8):
    per_rank, sums = [], np.zeros(len(queries), dtype=np.uint64)
    for rank in range(world if world <= 4 else 2):
        ctx.set_timing(1)
        best, steps = 1e9, 0
        for _ in range(reps):
            ctx.batch_join(rank, world)
            raw = ctx.batch_download()
            st = ctx.stats()
            best = min(best, st["last_join_ms"])
            steps = st["join_steps"]
        ctx.set_timing(0)
        sums += raw
        per_rank.append((round(best, 3), int(steps), round(st["join_steps"] / max(32 * st["join_warp_iters"], 1), 3)))
    if world <= 4:
        assert np.array_equal(sums, total), "the shares of the ranks do not add up"
    out[world] = per_rank
    print(world, per_rank, flush=True)
env = {k: v for k, v in os.environ.items() if k.startswith("GPE_")}
print(json.dumps(dict(workload=name, env=env, join_ms_steps_laneutil_per_rank=out)))
ctx.close()
