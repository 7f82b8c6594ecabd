#!/bin/bash
# One gpurun call: the GPU parity suite, then short benches.  usage: tools/gpu_check.sh TAG [workloads...]
TAG=${1:-x}; shift
WLS=${@:-config2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
for WL in $WLS; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $WL > gpurun_out/${TAG}_bench_${WL}.json 2> gpurun_out/${TAG}_bench_${WL}.err
  echo "bench $WL rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_${WL}.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","answers_checksum")}, d["e2e"]["value"], d["roofline"]["stage_ms_per_step"], d["oracle_parity"], d["build"]["build_table_ms"], d["build"]["set_graph_s"])
except Exception as e: print("no json", e)
PY
  tail -3 gpurun_out/${TAG}_bench_${WL}.err
done
if [ -n "$GPE_LAUNCH_LIST" ]; then
  for WL in $GPE_LAUNCH_LIST; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-streaming --workload $WL > gpurun_out/${TAG}_ncu_bench_${WL}.log 2>&1
    echo "ncu $WL rc=$?"
  done
fi
