"""GPU (needs >= 2 devices, skipped otherwise): 2-rank NCCL update equals the single-rank update on the concatenated batch."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
@pytest.mark.parametrize('math', ['fp32', 'tf32x3', 'bf16x3'])
def test_two_rank_nccl_update_equals_single_rank(math):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, PAACB_CHECK_MATH=math)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29541', os.path.join(ROOT, 'tools', 'multi_gpu_check.py')]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=540)
    assert r.returncode == 0 and 'multi_gpu_check ok' in r.stdout, r.stdout[-3000:]


@pytest.mark.timeout(900)
@pytest.mark.parametrize('flags', [[], ['--graphs', 'false', '--train_forward', 'batched']])
def test_two_rank_train_ends_with_identical_parameters(flags):
    """PAACLearner.train() under torchrun (cfg4's code path): both ranks end with the same bits."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29547', os.path.join(ROOT, 'tools', 'train_ranks_check.py')] + flags
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=840)
    assert r.returncode == 0 and 'train_ranks_check ok' in r.stdout, r.stdout[-3000:]
