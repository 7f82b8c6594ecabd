"""`host/main -g N -q DIR` on a config-5-style batch (mixed sparse/dense queries of 4-16 vertices, n=MAX) against `-g 1`:
the same `Answer Number` lines.  Writes the graphs in the reference's formats, runs the CLI twice, compares, keeps the log.
   usage: python tools/cli_multi_gpu_check.py N [workload] [out.log]"""
import os, subprocess, sys, tempfile, time
ROOT = os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from gnn_pe_b200 import graph_io

n_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 2
name = sys.argv[2] if len(sys.argv) > 2 else "config5_small"
log = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", f"cli_g{n_gpu}_{name}.log")
w, g, queries = bench.load_workload(name)
p = max(w["p"], n_gpu)
d = tempfile.mkdtemp(prefix="cli_multi_") + "/"
os.makedirs(d + "gnn-pe/partitions", exist_ok=True)
for i in range(p):
    os.makedirs(d + f"gnn-pe/partitions/partition-{i}", exist_ok=True)
os.makedirs(d + "queries", exist_ok=True)
graph_io.write_graph(d + "data.graph", g)
graph_io.write_membership(d + "gnn-pe/membership.txt", graph_io.degree_order(g), graph_io.block_membership(g.V, p))
for i, q in enumerate(queries):
    graph_io.write_graph(d + f"queries/q{i:04d}.graph", q)
exe = os.path.join(ROOT, "host", "main")
common = [exe, "-f", d, "-d", d + "data.graph", "-q", d + "queries", "-m", "online", "-p", str(p), "-l", str(w["l"]), "-e", str(w["e"])]
out = {}
for n in (1, n_gpu):
    t0 = time.time()
    out[n] = subprocess.check_output(common + ["-g", str(n)]).decode()
    print(f"-g {n}: {time.time() - t0:.1f} s wall;", out[n].strip().splitlines()[-1], flush=True)
ans = {n: [ln for ln in out[n].splitlines() if "Answer Number" in ln] for n in out}
assert len(ans[1]) == len(queries) and ans[1] == ans[n_gpu], "answers differ between -g 1 and -g %d" % n_gpu
open(log, "w").write(f"# host/main -q DIR on {name} ({len(queries)} queries), -g 1 and -g {n_gpu}: identical answers\n" +
                     "\n".join(f"## -g {n}\n" + out[n] for n in out))
print("identical answers:", len(ans[1]), "queries; total", sum(int(a.split("Answer Number: ")[1]) for a in ans[1]))
