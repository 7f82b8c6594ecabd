#!/bin/bash
# One gpurun call (1 GPU): parity suite, the GNN-PGE bench leg, the real-reference baseline.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --filter pge > gpurun_out/${TAG}_bench_pge_config2.json 2> gpurun_out/${TAG}_bench_pge.err
echo "pge bench rc=$?"; tail -c 1200 gpurun_out/${TAG}_bench_pge_config2.json; tail -3 gpurun_out/${TAG}_bench_pge.err
timeout 1500 python tools/real_reference_baseline.py --out=gpurun_out/${TAG}_cpu_baseline_real.json > gpurun_out/${TAG}_real_baseline.log 2>&1
echo "real baseline rc=$?"; tail -c 1500 gpurun_out/${TAG}_real_baseline.log
