#!/bin/bash
# ncu full capture of the join's persistent kernel on a workload (default config2), second step.  usage: tools/gpu_profile_dfs.sh TAG [workload]
TAG=${1:-x}; WL=${2:-config2}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k3_dfs_kernel --launch-skip 1 --launch-count 1 \
  -o gpurun_out/${TAG}_dfs -f python tools/profile_target.py $WL 2 0 > gpurun_out/${TAG}_ncu_dfs.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/${TAG}_ncu_dfs.log
