#!/bin/bash
# One gpurun --gpus N call: (N=2: the multi-GPU tests), then bench.py under torchrun for each workload.
# usage: tools/gpu_multi.sh TAG N [workloads...]
TAG=${1:-x}; N=${2:-2}; shift; shift
WLS=${@:-config2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${TAG}_gpu.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/${TAG}_pytest_multigpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_multigpu.log
  tail -5 gpurun_out/${TAG}_pytest_multigpu.log
fi
for WL in $WLS; do
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
     bench.py --gpus $N --steps 10 --warmup 3 --workload $WL --no-cpu-baseline > gpurun_out/${TAG}_bench_${WL}_${N}gpu.json 2> gpurun_out/${TAG}_bench_${WL}_${N}gpu.err
  echo "bench $WL x$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${WL}_${N}gpu.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","answers_checksum")}, "e2e", d["e2e"]["value"], d["roofline"]["stage_ms_per_step"], d["oracle_parity"], "build", d["build"]["build_table_ms"], d["build"]["set_graph_s"], d["build"]["setup_s"])
    print([ (r["rank"], r["stage_ms_per_step"]["join"], r["stage_ms_per_step"]["scan"]) for r in d["per_rank"]])
except Exception as e: print("no json", e)
PY
  tail -3 gpurun_out/${TAG}_bench_${WL}_${N}gpu.err
done
