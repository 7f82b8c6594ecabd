#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_config2.json 2> gpurun_out/${TAG}_bench_config2.err
echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_config2.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), d["roofline"]["frac"], d["roofline"]["ms_per_launch"], d["roofline"]["streaming"]["frac"], d["roofline"]["stage_ms_per_step"], d["oracle_parity"]["ok"], d["cpu_baseline"]["value"])
PY
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
