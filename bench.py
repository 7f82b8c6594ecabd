#!/usr/bin/env python
"""bench.py -- online queries/sec and dominance-scan GB/s on BASELINE.json's workload.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libgpe.so)
  python bench.py --impl reference --gpus N ...            # the reference algorithm on the host cores

A step = one batch of the workload's queries through the whole online stage (filter + merge + join).
`value`  : queries/s with the batch already uploaded (device-resident inputs), CUDA events, max over ranks.
`e2e`    : queries/s through the one-call C ABI entry gpe_query_batch with HOST buffers: host planning,
           H2D of the batch, kernels, D2H of the answers all inside the timed region.
`roofline`: the dominance-scan kernel as launched inside the timed steps (deferred CUDA events on the
           library's stream), plus a streaming pass (pruning off, every table row compared) afterwards.
Nothing here reads /root/reference.  oracle/ is used only for the cpu_baseline leg and --impl reference.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on, fits one B200
    "config2": dict(kind="chung_lu", V=1_000_000, E=10_000_000, labels=20, gamma=3.0, cap=1000, seed=2022,
                    l=2, e=2, n_queries=100, q_vertices=8, q_seed=2023, p=8,
                    desc="synthetic power-law 1M v / 10M e / 20 labels, l=2, e=2, 100 random-walk 8-vertex queries"),
    # scaled-down copy for development runs
    "small": dict(kind="chung_lu", V=100_000, E=1_000_000, labels=20, gamma=3.0, cap=1000, seed=2022,
                  l=2, e=2, n_queries=100, q_vertices=8, q_seed=2023, p=8,
                  desc="synthetic power-law 100K v / 1M e / 20 labels, l=2, e=2, 100 random-walk 8-vertex queries"),
    # BASELINE.json configs[2]: sharded over 2/4/8 GPUs.  Poisson degrees (mean 20): 2.0 B table rows = 144 GB of scan
    # tiles + 24 GB of ids, i.e. 84 / 42 / 21 GB per GPU at 2 / 4 / 8 (the power-law variant, 5.2 B rows, does not fit 2)
    "config3": dict(kind="uniform_native", V=10_000_000, E=100_000_000, labels=50, seed=2024,
                    l=2, e=2, n_queries=100, q_vertices=8, q_seed=2025, p=8, min_gpus=2,
                    desc="synthetic Poisson-degree 10M v / 100M e / 50 labels, p=8, l=2, e=2, 100 random-walk 8-vertex queries"),
    # BASELINE.json configs[4]: query-batch throughput on the config 3 graph, n = MAX answers
    "config5": dict(kind="uniform_native", V=10_000_000, E=100_000_000, labels=50, seed=2024,
                    l=2, e=2, n_queries=1000, q_vertices=(4, 16), q_mixed=True, q_seed=2026, p=8, min_gpus=2,
                    desc="synthetic Poisson-degree 10M v / 100M e / 50 labels, p=8, l=2, e=2, 1000 mixed sparse/dense "
                         "random-walk queries of 4-16 vertices, n=MAX"),
    # BASELINE.json configs[3]: longer paths.  Poisson degrees (mean 20), 20 labels: ~1.8 x 10^10 four-vertex paths; 160-byte
    # rows materialised would be 3.2 TB, so the table is held as vertex ids only (16 bytes per row: 320 GB, 42 GB per GPU
    # at 8, 80 at 4; 2 GPUs are not enough) and the scan gathers the rest (GPE_TABLE_IDS, chosen automatically).  Patched-oracle semantics for l=3
    "config4": dict(kind="uniform_native", V=5_000_000, E=50_000_000, labels=20, seed=2027,
                    l=3, e=4, n_queries=100, q_vertices=12, q_seed=2028, p=8, min_gpus=4,
                    desc="synthetic Poisson-degree 5M v / 50M e / 20 labels, p=8, l=3, e=4, 100 dense (induced) random-walk 12-vertex queries"),
    "config4_small": dict(kind="uniform_native", V=100_000, E=1_000_000, labels=20, seed=2027,
                          l=3, e=4, n_queries=50, q_vertices=12, q_seed=2028, p=8,
                          desc="synthetic Poisson-degree 100K v / 1M e / 20 labels, p=8, l=3, e=4, 50 dense (induced) random-walk 12-vertex queries"),
    # down-scaled copies of configs 3 / 5 (same generator and query mix) for development runs and the CPU tests
    "config3_small": dict(kind="uniform_native", V=200_000, E=2_000_000, labels=50, seed=2024,
                          l=2, e=2, n_queries=100, q_vertices=8, q_seed=2025, p=8,
                          desc="synthetic Poisson-degree 200K v / 2M e / 50 labels, p=8, l=2, e=2, 100 random-walk 8-vertex queries"),
    "config5_small": dict(kind="uniform_native", V=200_000, E=2_000_000, labels=50, seed=2024,
                          l=2, e=2, n_queries=200, q_vertices=(4, 16), q_mixed=True, q_seed=2026, p=8,
                          desc="synthetic Poisson-degree 200K v / 2M e / 50 labels, p=8, l=2, e=2, 200 mixed sparse/dense "
                               "random-walk queries of 4-16 vertices, n=MAX"),
}


def make_graph(w):
    from gnn_pe_b200 import synth
    if w["kind"] == "chung_lu":
        return synth.chung_lu_graph(w["V"], w["E"], w["labels"], w["gamma"], w["cap"], w["seed"])
    return synth.uniform_graph_native(w["V"], w["E"], w["labels"], w["seed"])


def make_queries(w, g):
    from gnn_pe_b200 import synth
    return synth.query_batch(g, w["n_queries"], w["q_vertices"], seed=w["q_seed"], mixed=w.get("q_mixed", False))


def load_workload(name, rank=0, world=1, barrier=None):
    """Rank 0 generates the graph once and leaves it as raw arrays in shared memory; the other ranks map it."""
    from gnn_pe_b200 import graph_io
    w = WORKLOADS[name]
    graph_key = f"{w['kind']}_{w['V']}_{w['E']}_{w['labels']}_{w['seed']}"
    base = os.environ.get("GPE_BENCH_CACHE", "/dev/shm/gpe_bench_cache" if os.path.isdir("/dev/shm") else "/tmp/gpe_bench_cache")
    paths = {k: os.path.join(base, f"{graph_key}.{k}.npy") for k in ("offsets", "nbrs", "labels")}
    if rank == 0 and not all(os.path.exists(p) for p in paths.values()):
        os.makedirs(base, exist_ok=True)
        g = make_graph(w)
        for k, p in paths.items():
            np.save(p + ".tmp.npy", getattr(g, k))
            os.replace(p + ".tmp.npy", p)
    if barrier:
        barrier()
    g = graph_io.CSRGraph(*(np.load(paths[k], mmap_mode="r") for k in ("offsets", "nbrs", "labels")))
    g = graph_io.CSRGraph(np.ascontiguousarray(g.offsets), np.ascontiguousarray(g.nbrs), np.ascontiguousarray(g.labels))
    return w, g, make_queries(w, g)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, f"/tmp/gpe_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device), "-lms", "100"], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for line in open(self.path):
            t = [x.strip() for x in line.split(",")]
            if len(t) < 8:
                continue
            try:
                sm.append(float(t[1]))
                mx = max(mx, float(t[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], t[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


def cpu_reference_leg(w, g, queries, budget_s, threads, expect=None):
    """The reference's algorithm on the host cores (oracle 'port': all-pairs leaf compare walked from the CSR,
    OpenMP over start vertices, then the reference's refinement).  Bounded sample of the workload's queries."""
    from gnn_pe_b200 import graph_io
    from oracle import oracle
    og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
    sorted_nodes = graph_io.degree_order(g)
    _, vde = og.embeddings(w["e"])
    times, answers = [], []
    t_start = time.time()
    for i, q in enumerate(queries):
        oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
        t0 = time.time()
        n, t3 = oracle.online_streaming(og, oq, w["l"] + 1, w["e"], sorted_nodes, vde, threads=threads)
        times.append(time.time() - t0)
        answers.append(n)
        if time.time() - t_start > budget_s:
            break
    ok = None
    if expect is not None:
        ok = all(int(a) == int(b) for a, b in zip(answers, expect))
    return dict(n=len(times), seconds=float(sum(times)), answers=answers, parity_ok=ok)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w, g, queries = load_workload(args.workload)
    threads = os.cpu_count() or 1
    per_step_budget = max(5.0, 150.0 / max(args.steps + args.warmup, 1))
    for _ in range(args.warmup):
        cpu_reference_leg(w, g, queries[:1], 0.0, threads)
    tot_q, tot_s = 0, 0.0
    for _ in range(args.steps):
        r = cpu_reference_leg(w, g, queries, per_step_budget, threads)
        tot_q += r["n"]
        tot_s += r["seconds"]
    qps = tot_q / tot_s
    sample = f"first {tot_q // max(args.steps, 1)} of {len(queries)} queries per step, full data graph"
    out = dict(impl="reference", metric="online queries/sec", value=qps, unit="queries/s", n_gpus=args.gpus,
               steps=args.steps, warmup=args.warmup, ms_per_step=1000.0 * tot_s / max(args.steps, 1),
               higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
               config=dict(workload=w["desc"], name=args.workload),
               cpu_baseline=dict(value=qps, unit="queries/s", cores=threads, kind="port", sample=sample,
                                 note="oracle/_ref (the real binary) cannot build its R*-tree at this size "
                                      "(~100 us/row => hours); this is the restatement of its leaf compare + refinement"),
               e2e=dict(value=qps, unit="queries/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))
    return 0


def run_pge(args):
    """The GNN-PGE variant of the filter on one GPU: per-vertex path groups built on the device, the
    label classes the batch asks for scanned (k4_pge_scan), then the shared compaction / order / join.  Same metric."""
    import torch
    from gnn_pe_b200 import gpe
    from oracle import oracle
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise SystemExit("--filter pge runs on one GPU (one row per data vertex: the table is small; replicas only)")
    w, g, queries = load_workload(args.workload)
    e, pl = w["e"], 2  # GNN-PGE's default path length: 2 vertices per path
    ctx = gpe.GpeContext(0)
    x, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, e)
    ctx.set_graph(g.offsets, g.nbrs, g.labels)
    ctx.set_embeddings(vde)
    t0 = time.perf_counter()
    ctx.pge_build(pl, x)
    build_ms = (time.perf_counter() - t0) * 1e3
    nq = len(queries)
    ctx.pge_batch_upload(queries)
    for _ in range(max(args.warmup, 3)):
        ctx.pge_batch_filter()
        ctx.batch_join()
    answers = ctx.batch_download()
    ctx.collect_timings()
    ctx.set_timing(2)
    stream = torch.cuda.ExternalStream(ctx.stream)
    sampler = ClockSampler(0)
    sampler.start()
    with torch.cuda.stream(stream):
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            ctx.pge_batch_filter()
            ctx.batch_join()
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
    stages = ctx.collect_timings()
    ctx.set_timing(0)
    assert np.array_equal(answers, ctx.batch_download())
    st = ctx.stats()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        a3 = ctx.pge_query_batch(queries)
    e2e_s = time.perf_counter() - t1
    assert np.array_equal(a3, np.minimum(answers, gpe.LIMIT_MAX))
    s3 = ctx.stats()
    clocks = sampler.stop()
    peak, peak_src = peaks()
    pde = pl * e
    row_bytes = 3 * pde * 8 + 4  # what the leaf test reads: label box lo/hi, upper corner of the embedding box, degree
    scan_ms = stages["scan"]["ms"] / max(stages["scan"]["launches"], 1)
    scan_bytes = st["scan_rows"] * row_bytes
    cpu = None
    if not args.no_cpu_baseline:
        og = oracle.OracleGraph.from_csr(g.offsets, g.nbrs, g.labels)
        t0, n, ok = time.time(), 0, True
        oracle.pge_groups(og, pl, e)
        t_groups = time.time() - t0
        t0 = time.time()
        for i, q in enumerate(queries):
            oq = oracle.OracleGraph.from_csr(q.offsets, q.nbrs, q.labels)
            ok = ok and oracle.pge_online(og, oq, pl, e) == int(min(answers[i], gpe.LIMIT_MAX))
            n += 1
            if time.time() - t0 > args.cpu_baseline_seconds:
                break
        cpu = dict(value=n / (time.time() - t0), unit="queries/s", cores=1, kind="port",
                   sample=f"first {n} of {nq} queries (orc_pge_online: groups recomputed per call), data-side groups {t_groups:.1f} s",
                   parity_with_gpu_answers=ok)
    out = dict(metric="online queries/sec", value=nq * args.steps / (ms / 1e3), unit="queries/s", n_gpus=1, steps=args.steps,
               warmup=max(args.warmup, 3), ms_per_step=ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
               dtype="f64", data="synthetic", filter="gnn-pge",
               config=dict(workload=w["desc"], name=args.workload, filter=f"GNN-PGE path groups, pl={pl}, e={e}"),
               gpu_launches=int(st["kernel_launches"]), clocks=clocks,
               e2e=dict(value=nq * args.steps / e2e_s, unit="queries/s", h2d_bytes_per_step=int(s3["h2d_bytes"]),
                        d2h_bytes_per_step=int(s3["d2h_bytes"]), ms_per_step=1e3 * e2e_s / args.steps, api="gpe_pge_query_batch"),
               roofline=dict(bound="hbm", kernel=f"k4_pge_scan_kernel<{pde}>", achieved=scan_bytes / max(scan_ms, 1e-9) / 1e6, peak=peak,
                             unit="GB/s", frac=scan_bytes / max(scan_ms, 1e-9) / 1e6 / peak, traffic=None, peak_source=peak_src,
                             algorithmic_bytes_per_launch=int(scan_bytes), rows_per_launch=int(st["scan_rows"]), row_bytes=row_bytes,
                             ms_per_launch=scan_ms, stage_ms_per_step={k: v["ms"] / args.steps for k, v in stages.items()},
                             note="one row per data vertex of the label classes the batch asks for; at 1 M vertices the launch is "
                                  "tens of microseconds: latency, not bandwidth"),
               cpu_baseline=cpu, build=dict(pge_build_ms=build_ms), answers_checksum=int(np.minimum(answers, gpe.LIMIT_MAX).sum()),
               n_candidates=int(st["n_candidates"]))
    print(json.dumps(out))
    ctx.close()
    return 0


def expected_answers(name):
    """Oracle answers of sampled queries (tests/golden/config_answers.json, made by tests/golden/make_config_answers.py)."""
    path = os.path.join(ROOT, "tests", "golden", "config_answers.json")
    try:
        rec = json.load(open(path)).get(name)
    except Exception:
        rec = None
    return {int(k): int(v) for k, v in rec["answers"].items()} if rec else {}


def static_json(*parts):
    try:
        return json.load(open(os.path.join(ROOT, *parts)))
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("GPE_BENCH_WORKLOAD", "config2"), choices=list(WORKLOADS))
    ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-streaming", action="store_true")
    ap.add_argument("--table-layout", type=int, default=0, choices=[0, 1, 2],
                    help="0 auto (rows while they fit), 1 materialised rows, 2 vertex ids only (rows gathered by the scan)")
    ap.add_argument("--filter", default="path", choices=["path", "pge"],
                    help="path: GNN-PE's path-table dominance scan (the headline); pge: the GNN-PGE per-vertex path-group filter")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    if args.filter == "pge":
        return run_pge(args)

    import torch  # before libgpe: the process then carries ONE NCCL (PyTorch's bundled copy), which libgpe binds at run time
    import torch.distributed as dist
    from gnn_pe_b200 import gpe, graph_io

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: libgpe has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # torch.distributed only carries the control plane (rendezvous, barriers, the max over ranks of the timings);
        # the data path -- candidate all-gather, count all-reduce -- is NCCL inside libgpe (gpe_comm_init)
        dist.init_process_group("gloo")
    barrier = (lambda: dist.barrier()) if world > 1 else None

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w, g, queries = load_workload(args.workload, rank, world, barrier)
    if world < w.get("min_gpus", 1):
        raise SystemExit(f"workload {args.workload} needs at least {w['min_gpus']} GPUs (table size); got {world}")
    L, e, p = w["l"] + 1, w["e"], max(w["p"], world)
    ctx = gpe.GpeContext(local)
    nccl_version = 0
    if world > 1:
        ids = [gpe.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(rank, world, ids[0])
        nccl_version = ctx.comm_info()[2]
    t0 = time.time()
    _, vde = gpe.host_gen_vde(g.offsets, g.nbrs, g.labels, e)
    sorted_nodes = graph_io.degree_order(g)
    membership = graph_io.block_membership(g.V, p)
    t1 = time.time()
    ctx.set_graph(g.offsets, g.nbrs, g.labels)
    ctx.sync()
    t_graph = time.time() - t1
    ctx.set_embeddings(vde)
    ctx.set_timing(1)
    n_rows, rows_pp = ctx.enumerate(L, sorted_nodes, membership, p)
    ctx.set_table_layout(args.table_layout)
    table_rows = ctx.build_table_shard()
    st = ctx.stats()
    row_bytes = st["row_bytes"]
    peak, peak_src = peaks()
    # K1 (offline table build) against the HBM roofline, SURVEY.md 8(d): N x (4L + row_bytes) written + the CSR read once
    k1_bytes = table_rows * st["stored_row_bytes"] + 4 * (g.V + 1 + 2 * g.E + g.V)
    build = dict(enumerate_ms=st["last_enumerate_ms"], build_table_ms=st["last_build_ms"], rows=n_rows,
                 table_rows_this_rank=table_rows, table_gb_this_rank=table_rows * st["stored_row_bytes"] / 1e9,
                 table_layout="ids only (rows gathered by the scan)" if st["table_ids_only"] else "materialised rows + ids",
                 stored_row_bytes=int(st["stored_row_bytes"]),
                 set_graph_s=t_graph, setup_s=time.time() - t0,
                 roofline_k1=dict(bound="hbm", kernels="k1_hist + k1_fill + k1_expand (whole table build)",
                                  algorithmic_bytes=int(k1_bytes), ms=st["last_build_ms"],
                                  achieved=k1_bytes / max(st["last_build_ms"], 1e-9) / 1e6, peak=peak, unit="GB/s",
                                  frac=k1_bytes / max(st["last_build_ms"], 1e-9) / 1e6 / peak,
                                  note="kernel time of the build (histogram + scan, fill, expand); the cudaMalloc of the table is not included; per-kernel times in profiles/"))
    ctx.set_timing(0)

    stream = torch.cuda.ExternalStream(ctx.stream)
    nq = len(queries)
    limits = [gpe.LIMIT_MAX] * nq

    with torch.cuda.stream(stream):
        ctx.batch_upload(queries, limits)
        for _ in range(args.warmup):
            ctx.batch_step()
        answers = ctx.batch_finish()
        launches0 = ctx.stats()["kernel_launches"]
        ctx.collect_timings()
        ctx.set_timing(2)
        sampler = ClockSampler(local)
        sampler.start()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            ctx.batch_step()
        ev1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        stages = ctx.collect_timings()
        ctx.set_timing(0)
        answers2 = ctx.batch_finish()
        st_step = ctx.stats()
        launches = st_step["kernel_launches"] - launches0
        assert np.array_equal(answers, answers2)

        # ---- e2e: host buffers in, answers out.  ONE C-ABI call (gpe_query_batches) for `steps` batches of the workload:
        # per batch the host planning (dfs_query, gen_vde(query), gen_query_pde -- the span the reference itself times,
        # main.cpp:148-152), the H2D copy of the batch, the kernels, the exchange (N > 1: NCCL all-gather + all-reduce
        # inside the library) and the D2H copy of the answers are all inside the timed region; the library plans batch
        # i+1 while the GPU works on batch i.  `e2e_single_call` is the same through one gpe_query_batch call per step
        # (no overlap), 1 GPU only.
        prepared = ctx.prepare_batches([queries] * args.steps, [limits] * args.steps)
        warm = ctx.prepare_batches([queries] * 2, [limits] * 2)
        ctx.run_batches(warm)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        s2 = ctx.stats()
        t1 = time.perf_counter()
        outs = ctx.run_batches(prepared)
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t1)
        assert all(np.array_equal(a3, answers) for a3 in outs)
        s3 = ctx.stats()
        e2e = dict(value=nq * args.steps / e2e_s, unit="queries/s", h2d_bytes_per_step=int(s3["h2d_bytes"] // args.steps),
                   d2h_bytes_per_step=int(s3["d2h_bytes"] // args.steps), ms_per_step=1000.0 * e2e_s / args.steps,
                   launches_per_step=(s3["kernel_launches"] - s2["kernel_launches"]) / args.steps,
                   join_reruns=int(s3["join_reruns"] - s2["join_reruns"]), exchange_redos=int(s3["exchange_redos"] - s2["exchange_redos"]),
                   n_candidates=int(s3["n_candidates"]), exchange_bytes=int(s3["exchange_bytes"]),
                   host_ms_per_step=dict(upload=s3["host_upload_ms"] / args.steps, enqueue=s3["host_enqueue_ms"] / args.steps,
                                         plan_next=s3["host_plan_ms"] / args.steps, finish_wait=s3["host_finish_ms"] / args.steps),
                   api=f"gpe_query_batches: {args.steps} batches in one call (host plan of batch i+1 overlapped with the GPU work of batch i; "
                       "per batch: plan + H2D + kernels" + (" + NCCL all-gather + all-reduce" if world > 1 else "") + " + D2H)")
        if world == 1:
            ctx.query_batch(queries, limits)
            t1 = time.perf_counter()
            for _ in range(args.steps):
                a4 = ctx.query_batch(queries, limits)
            e2e1_s = time.perf_counter() - t1
            assert np.array_equal(a4, answers)
            e2e["single_call_per_step"] = dict(value=nq * args.steps / e2e1_s, ms_per_step=1000.0 * e2e1_s / args.steps,
                                               api="gpe_query_batch, one call per step, nothing overlapped")
        clocks = sampler.stop()

        # ---- streaming pass: pruning off, every row of this rank's table against one query's plan paths ----
        streaming = None
        if not args.no_streaming:
            q0 = queries[0]
            plan = gpe.host_query_plan(q0.offsets, q0.nbrs, q0.labels, L, e)
            ctx.set_timing(1)
            best, tot_ms, reps = 1e30, 0.0, 5
            for i in range(reps + 1):
                ctx.filter(plan, q0.V, gpe.FILTER_NO_PRUNE)
                s = ctx.stats()
                if i:
                    best = min(best, s["last_scan_ms"])
                    tot_ms += s["last_scan_ms"]
            ctx.set_timing(0)
            sbytes = s["scan_rows"] * row_bytes
            streaming = dict(rows=int(s["scan_rows"]), bytes=int(sbytes), ms_avg=tot_ms / reps, ms_best=best,
                             achieved=sbytes / (tot_ms / reps) / 1e6, unit="GB/s", frac=sbytes / (tot_ms / reps) / 1e6 / peak,
                             plan_paths=int(len(plan["vids"])), rank=rank)

    scan = stages["scan"]
    scan_ms = scan["ms"] / max(scan["launches"], 1)
    scan_bytes = st_step["scan_rows"] * row_bytes
    achieved = scan_bytes / scan_ms / 1e6 if scan_ms > 0 else 0.0
    # physical DRAM bytes of the in-step scan launch: from the kept ncu capture of the SAME workload at 1 GPU, else null
    traffic, traffic_src = None, None
    tr = static_json("profiles", "scan_traffic.json")
    if tr and world == 1 and tr.get("workload", "config2") == args.workload:
        traffic, traffic_src = tr.get("in_step_dram_bytes_per_launch"), "profiles/scan_traffic.json (ncu --set full capture, not re-measured by this run)"
    roofline = dict(bound="hbm", kernel=(f"k2_scan_ids_kernel<{L},{e}> (ids-only table: {4 * L} B/row streamed, the rest gathered)"
                                         if st_step["table_ids_only"] else f"k2_scan_kernel<{L},{e}>") + " (in-step, label-bucketed work list)",
                    achieved=achieved, peak=peak,
                    unit="GB/s", frac=achieved / peak, traffic=traffic, traffic_source=traffic_src, peak_source=peak_src,
                    algorithmic_bytes_per_launch=int(scan_bytes), rows_per_launch=int(st_step["scan_rows"]),
                    row_bytes=int(row_bytes), ms_per_launch=scan_ms,
                    tiles_examined=int(st_step["scan_items"]), tiles_unpruned=int(st_step["scan_items_unpruned"]),
                    share_of_step=scan["ms"] / max(stages["scan"]["ms"] + stages["select"]["ms"] + stages["compact"]["ms"] + stages["join"]["ms"], 1e-9),
                    streaming=streaming, rank=rank,
                    stage_ms_per_step={k: v["ms"] / args.steps for k, v in stages.items()})
    # the join is latency / divergence bound: candidate tests per second and what the kept ncu capture says about it
    join_ms = stages["join"]["ms"] / args.steps
    join = dict(ms_per_step=join_ms, candidate_tests_per_step=int(st_step["join_steps"]),
                tests_per_s=st_step["join_steps"] / max(join_ms, 1e-9) * 1e3,
                lane_utilisation=st_step["join_steps"] / max(32 * st_step["join_warp_iters"], 1),
                exports=int(st_step["join_exports"]), donations=int(st_step["join_donations"]), rank=rank,
                ncu=static_json("profiles", "join_ncu.json"))

    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        r = cpu_reference_leg(w, g, queries, args.cpu_baseline_seconds, threads, expect=answers)
        cpu_baseline = dict(value=r["n"] / r["seconds"], unit="queries/s", cores=threads, kind="port",
                            what="brute-force CPU port of the leaf compare (custom.h:407-435, all pairs, no index) + the "
                                 "reference's refinement; the real binary's indexed path is in cpu_baseline_real",
                            sample=f"first {r['n']} of {nq} queries, full data graph, {r['seconds']:.1f} s",
                            parity_with_gpu_answers=r["parity_ok"])
    per_rank = None
    if world > 1:
        mine = dict(rank=rank, table_rows=int(table_rows), stage_ms_per_step=roofline["stage_ms_per_step"],
                    join_tests=int(st_step["join_steps"]), scan_rows=int(st_step["scan_rows"]))
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = gathered
        dist.barrier()

    if rank == 0:
        exp = expected_answers(args.workload)
        bad = {i: (int(answers[i]), v) for i, v in exp.items() if i < nq and int(answers[i]) != min(v, gpe.LIMIT_MAX)}
        out = dict(metric="online queries/sec", value=nq * args.steps / (ms / 1000.0), unit="queries/s", n_gpus=world,
                   steps=args.steps, warmup=args.warmup, ms_per_step=ms / args.steps, higher_is_better=True,
                   scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                   config=dict(workload=w["desc"], name=args.workload, partitions=p,
                               parallelism=f"path table sharded over {world} GPU(s) by partition, NCCL inside libgpe" if world > 1 else "1 GPU",
                               l2_policy=f"scan reads {scan_bytes / 1e6:.0f} MB per step from a "
                                         f"{table_rows * st_step['stored_row_bytes'] / 1e9:.1f} GB table (> 126 MB L2), no flush"),
                   gpu_launches=int(launches), clocks=clocks, e2e=e2e, roofline=roofline, join=join, cpu_baseline=cpu_baseline,
                   cpu_baseline_real=static_json("profiles", "cpu_baseline_real.json"),
                   per_rank=per_rank, build=build, nccl=dict(version=nccl_version, ranks=world, via="libgpe gpe_comm_init (ncclCommInitRank)") if world > 1 else None,
                   answers_checksum=int(answers.sum()), answers_nonzero=int((answers > 0).sum()),
                   oracle_parity=dict(checked=len([i for i in exp if i < nq]), mismatches=bad, ok=not bad,
                                      source="tests/golden/config_answers.json (oracle.online_streaming on the CPU)"))
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    try:
        rc = main()
    except BaseException:  # noqa: BLE001
        # A rank that fails must not linger: its peers are (or will be) waiting inside a collective, and destructors that
        # synchronise with the GPU would wait for them in turn.  Print, then leave without running any.
        import traceback
        traceback.print_exc()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)
    sys.exit(rc)
