"""Run under torchrun with G ranks (one per GPU): the G-rank engine update (envs sharded, NCCL allreduce of the flat gradient,
1/G folded into paacb_clip_rmsprop) must equal the 1-rank update on the concatenated batch, and leave all ranks identical.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from paac_b200 import parallel
from paac_b200.engine import RolloutEngine
from paac_b200.policy_v_network import NaturePolicyVNetwork, NIPSPolicyVNetwork

ARCH = os.environ.get('PAACB_CHECK_ARCH', 'NATURE').upper()


def make(local, math):
    conf = dict(name='local_learning', num_actions=6, clip_norm=3.0, clip_norm_type='global', device='/gpu:%d' % local,
                entropy_regularisation_strength=0.02, seed=3, math=math)
    return (NIPSPolicyVNetwork if ARCH == 'NIPS' else NaturePolicyVNetwork)(conf)


def fill(eng, states, actions, values, rewards, over, lo, hi):
    eng.set_states(torch.from_numpy(np.ascontiguousarray(states[:, lo:hi])).cuda())
    eng.actions.copy_(torch.from_numpy(actions[:, lo:hi]).cuda())
    eng.values.copy_(torch.from_numpy(values[:, lo:hi]).cuda())
    eng.rewards.copy_(torch.from_numpy(rewards[:, lo:hi]).cuda())
    eng.over.copy_(torch.from_numpy(over[:, lo:hi]).cuda())


def main():
    rank, world, local = parallel.init_from_env('nccl')
    math = os.environ.get('PAACB_CHECK_MATH', 'fp32')
    N, T, A = 16 * world, 5, 6
    rng = np.random.RandomState(0)
    states = rng.randint(0, 256, (T + 1, N, 84, 84, 4)).astype(np.uint8)
    actions = rng.randint(0, A, (T, N)).astype(np.int32)
    values = rng.randn(T, N).astype(np.float32)
    rewards = rng.choice([-1.0, 0.0, 1.0], size=(T, N)).astype(np.float32)
    over = (rng.random_sample((T, N)) < 0.1).astype(np.float32)

    net = make(local, math)
    lo, hi = parallel.shard_range(rank, world, N)
    eng = RolloutEngine(net, hi - lo, T, world_size=world)
    fill(eng, states, actions, values, rewards, over, lo, hi)
    p0 = net.get_params()
    eng.update(0.0224)
    torch.cuda.synchronize()
    mine = net.params.clone()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    torch.distributed.all_gather(gathered, mine)
    for g in gathered:
        assert torch.equal(g, gathered[0]), 'ranks diverged'
    if rank == 0:
        ref = make(local, math)
        eng1 = RolloutEngine(ref, N, T, world_size=1)
        fill(eng1, states, actions, values, rewards, over, 0, N)
        eng1.update(0.0224)
        torch.cuda.synchronize()
        d_multi = net.get_params() - p0
        d_one = ref.get_params() - p0
        err = np.max(np.abs(d_multi - d_one)) / np.max(np.abs(d_one))
        nerr = abs(eng.norm.item() - eng1.norm.item()) / eng1.norm.item()
        print('multi_gpu_check world=%d arch=%s math=%s: delta err %.2e, norm err %.2e' % (world, ARCH, math, err, nerr), flush=True)
        assert err <= 1e-5 and nerr <= 1e-5      # measured: 4e-7 (one ulp of a parameter) / 2e-7
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
    if rank == 0:
        print('multi_gpu_check ok', flush=True)


if __name__ == '__main__':
    main()
