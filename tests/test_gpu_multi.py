"""Multi-GPU check (needs >= 2 GPUs; skipped on a single-GPU box): sharded run == unsharded run
bit for bit, and the NCCL moments/histogram all-reduce equals the statistics of the whole."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs at least 2 GPUs')
def test_sharded_run_is_bit_identical_to_single_gpu():
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}',
           '--master-addr', '127.0.0.1', '--master-port', '29611',
           os.path.join(ROOT, 'tests', 'multi_gpu_check.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f'MULTI_GPU_OK world={n}' in out.stdout
