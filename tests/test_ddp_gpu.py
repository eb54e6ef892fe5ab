"""Multi-rank GPU parity of the data-parallel path: two ranks on two B200s, different data per rank.
The engine's own bucket all-reduce (sd2_ddp_allreduce_bucket: NCCL through the C ABI), a torch DistributedDataParallel
wrapper with torch AdamW, the wrapper with FusedAdamW, and microbatch accumulation under no_sync() must all leave the
explicit average of the per-rank gradients on every rank.  Skipped on boxes with fewer than two GPUs (the driver's GPU
test box has one; run it with `gpurun --gpus 2 -- python -m pytest tests/test_ddp_gpu.py -m gpu`)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_gradient_averaging_routes_agree():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'tools', 'ddp_wrapper_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'DDP ROUTES OK' in r.stdout, r.stdout[-3000:]
