"""Data-parallel training against the single-GPU step on the concatenated batch, on 2 GPUs of one box (NCCL / symmetric
memory over NVLink).  Needs >= 2 visible GPUs (`gpurun --gpus 2`); skipped on a one-GPU box.  The checks live in
tests/dp_worker.py (one process per GPU under torch.distributed.run)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_step_equals_single_gpu_step_on_the_concatenated_batch():
    port = 29500 + os.getpid() % 400
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    assert r.stdout.count(" OK, ") == 6
