#!/bin/bash
# One gpurun call: the GPU parity suite, then a short bench of config 2.  usage: tools/gpu_check.sh TAG [pytest args]
TAG=${1:-x}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q "$@" > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_config2.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench_config2.json
tail -5 gpurun_out/${TAG}_bench.err
