"""GPU, world_size >= 2: the NVLink peer-memory gradient exchange against NCCL (skipped on single-GPU boxes)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_peer_allreduce_matches_nccl():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    world = 2
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'multi_gpu_allreduce.py')]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=280)
    assert res.returncode == 0, res.stdout[-3000:] + '\n' + res.stderr[-3000:]
    assert '"ok": true' in res.stdout


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_data_parallel_head_step_matches_reference_order():
    """bags_head_loss / GraphedHeadStep with grad_bucket= (exchange overlapped by dX) against local backward + NCCL mean."""
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tests', 'multi_gpu_head_step.py')]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=280)
    assert res.returncode == 0, res.stdout[-3000:] + '\n' + res.stderr[-3000:]
    assert '"ok": true' in res.stdout
