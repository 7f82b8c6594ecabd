#!/bin/bash
# tools/gpu_split_ab.sh TAG workload "ENV=.." ...   : tools/join_split_bench.py under different switches ("-" = defaults)
TAG=$1; WL=$2; shift; shift
mkdir -p gpurun_out
i=0
for V in "$@"; do
  i=$((i+1)); if [ "$V" = "-" ]; then V=""; fi
  echo "== [$V]"
  env $V timeout 600 python tools/join_split_bench.py $WL 5 2> gpurun_out/${TAG}_split${i}.err | tee gpurun_out/${TAG}_split${i}.log | grep -v "^{"
done
