"""Two-GPU data-parallel training step over NCCL; skipped on a single-GPU box.  The
world-size-2 host logic is covered on the CPU by tests/test_host.py (gloo)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_training_step_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tools", "bench_train_dist.py"),
           "--steps", "2", "--kinds", "planar", "nsf_h128"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 2
    for l in lines:
        assert l["n_gpus"] == 2 and l["allreduce_bytes"] == 4 * l["grad_elements"]
        assert l["replicas_bit_identical_after_steps"] is True
        # the averaged sharded gradient is the gradient of one process on the concatenated batch
        # (fp32 summation order / atomics differ between the two evaluations)
        assert l["sharded_vs_single_process_gradient_rel_err"] <= 1e-4, l
