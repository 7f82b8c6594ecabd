#!/bin/bash
# ncu --set full capture of one kernel (regex) of tools/profile_target.py.  usage: tools/gpu_ncu_kernel.sh TAG regex [workload] [skip]
TAG=$1; RX=$2; WL=${3:-config2}; SKIP=${4:-1}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:$RX --launch-skip $SKIP --launch-count 1 \
  -o gpurun_out/${TAG} -f python tools/profile_target.py $WL 2 0 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log
