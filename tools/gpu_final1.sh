#!/bin/bash
# Final 1-GPU evidence of a round: parity suite, the default bench line, launch list, ncu captures of the scan and the join.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_config2.json 2> gpurun_out/${TAG}_bench_config2.err
echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_config2.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
echo "reference rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_config2.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k2_scan_kernel --launch-skip 1 --launch-count 1 \
  -o gpurun_out/${TAG}_scan -f python tools/profile_target.py config2 2 0 > gpurun_out/${TAG}_ncu_scan.log 2>&1
echo "ncu scan rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k3_dfs_kernel --launch-skip 1 --launch-count 1 \
  -o gpurun_out/${TAG}_dfs -f python tools/profile_target.py config2 2 0 > gpurun_out/${TAG}_ncu_dfs.log 2>&1
echo "ncu dfs rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -c 200 gpurun_out/${TAG}_smoke.log
