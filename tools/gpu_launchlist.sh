#!/bin/bash
# ncu launch list of tools/profile_target.py (2 online steps, no streaming scans).  usage: tools/gpu_launchlist.sh TAG workload
TAG=$1; WL=${2:-config2}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_${WL}.csv \
   python tools/profile_target.py $WL 3 0 > gpurun_out/${TAG}_ncu_${WL}.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/${TAG}_ncu_${WL}.log
python tools/launch_summary.py gpurun_out/${TAG}_launches_${WL}.csv | grep -v "cub::\|k0_"
