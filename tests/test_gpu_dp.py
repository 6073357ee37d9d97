"""Data-parallel equality on real GPUs (last row; section 8e): two ranks under torch.distributed.run,
NCCL all-reduce of the flat gradient bucket.  Skipped on a box with fewer than two GPUs (the CPU / gloo coverage of the
host logic is tests/test_dp_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_step_equals_the_mean_of_the_oracle_shards():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dp_worker.py")]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    out = os.path.join(ROOT, "gpurun_out", "r2_dp_equality.jsonl")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert len(lines) == 4 and all(l["ok"] for l in lines), lines
