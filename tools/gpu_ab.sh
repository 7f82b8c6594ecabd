#!/bin/bash
# A/B of environment switches on one box: tools/gpu_ab.sh TAG workload "ENV1=a ENV2=b" "ENV1=c" ...   ("-" = defaults)
TAG=$1; WL=$2; shift; shift
mkdir -p gpurun_out
i=0
for V in "$@"; do
  i=$((i+1))
  if [ "$V" = "-" ]; then V=""; fi
  env $V timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-streaming --workload $WL > gpurun_out/${TAG}_ab${i}.json 2> gpurun_out/${TAG}_ab${i}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_ab${i}.json").read().strip().splitlines()[-1])
    print("[$V]", round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["stage_ms_per_step"], round(d["join"]["lane_utilisation"],3), d["build"]["build_table_ms"])
except Exception as e: print("[$V] failed", e)
PY
done
