"""N=2 over NCCL (needs 2 GPUs: run with gpurun --gpus 2).  Skipped on a 1-GPU box."""
import json
import os
import subprocess
import sys

import pytest

from tests.golden_util import ROOT

pytestmark = pytest.mark.gpu


def test_bench_two_gpus_matches_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, GPE_BENCH_WORKLOAD="small")
    one = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3",
                                   "--no-cpu-baseline"], env=env).decode().strip().splitlines()[-1]
    two = subprocess.check_output([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                                   "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "bench.py"),
                                   "--gpus", "2", "--steps", "2", "--warmup", "3", "--no-cpu-baseline"],
                                  env=env).decode().strip().splitlines()[-1]
    a, b = json.loads(one), json.loads(two)
    assert a["answers_checksum"] == b["answers_checksum"] and a["answers_nonzero"] == b["answers_nonzero"]
    assert b["n_gpus"] == 2
