"""Multi-GPU paths on a box with >= 2 B200s (skipped otherwise): the forward shards by site
batch with no collective (bench.py under torchrun), call_freq adds one NCCL exchange and must
stay bit-exact."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, script, *args, port=29533):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:]
    return json.loads(lines[-1])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_freq_exchange_over_nccl_is_bit_exact():
    out = _torchrun(2, "tools/freq_multigpu_check.py", "--records", "2000000")
    assert out["ok"] and out["world"] == 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bench_two_ranks_weak_scaling_line():
    out = _torchrun(2, "bench.py", "--gpus", "2", "--steps", "20", "--warmup", "3", port=29534)
    assert out["n_gpus"] == 2 and out["scaling"] == "weak" and out["value"] > 0 and out["gpu_launches"] > 0
