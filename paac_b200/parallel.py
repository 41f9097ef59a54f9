"""Multi-GPU plumbing: one process per GPU (torchrun), environments sharded over ranks, ONE collective per update.

The reference is single-device.  The B200 build shards the `emulator_counts` environments over the G
ranks of one box: rank r owns environments [r*N/G, (r+1)*N/G), its own Runners / workers / rollout buffers / acting
forwards.  Parameters and RMSProp slots are replicated.  Per update each rank computes the gradient of ITS batch mean
(paacb_backward), the flat fp32 gradient buffer is all-reduced (SUM) over NCCL/NVLink in one call, and
paacb_clip_rmsprop applies grad_scale = 1/G BEFORE the global norm, so that

    clip(mean_over_ranks(g_r)) == clip(gradient of the mean over the concatenated batch)

i.e. the G-rank update equals the single-learner update on all N environments (equal per-rank batches; the loss is a
batch mean, policy_v_network.py:49-53).  Every rank applies the identical update, so no parameter broadcast is needed
after step 0.
"""
import os

import torch


def shard_range(rank, world_size, n_items):
    """[first, last) of the contiguous slice rank `rank` owns; n_items must divide evenly (runners.py:17-18 has the
    same divisibility requirement for workers)."""
    if n_items % world_size != 0:
        raise ValueError('%d environments cannot be split evenly over %d ranks' % (n_items, world_size))
    per = n_items // world_size
    return rank * per, (rank + 1) * per


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's environment (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world_size, local_rank).  A no-op for single-process runs."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not torch.distributed.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
        else:
            torch.distributed.init_process_group(backend)
    return rank, world, local


def allreduce_mean_grads(flat_grads, world_size, group=None):
    """SUM-allreduce in place; the 1/world_size is NOT applied here (it is folded into the optimizer kernel)."""
    if world_size > 1:
        torch.distributed.all_reduce(flat_grads, op=torch.distributed.ReduceOp.SUM, group=group)
    return flat_grads
