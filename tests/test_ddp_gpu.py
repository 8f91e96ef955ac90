"""Multi-GPU correctness of the GAN step, run when the box has >= 2 GPUs (`gpurun --gpus 2`):
torchrun world 2, each rank takes its shard of the batch; gradients are summed through libfsr_b200's own NCCL
communicator (fsr_nccl_allreduce, captured inside the step's CUDA graph).  Asserted by tests/diag/ddp_check.py:
replicas bit-identical after 4 steps (2 eager + graph), parameters within Adam's sign-flip bound of the full batch
on one GPU, and the same through the torch.distributed fallback (FSR_NCCL_CAPI=0)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("capi", ["1", "0"])
def test_two_gpu_train_step_matches_single_gpu(capi):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, FSR_NCCL_CAPI=capi)
    port = 29500 + (os.getpid() % 400) + (0 if capi == "1" else 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "diag", "ddp_check.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DDP CHECK worst" in r.stdout
