"""Run under torchrun with G ranks: PAACLearner.train() on the synthetic game (cfg4's code path: environments sharded over
the ranks, actor_learner.py:38-45 + train.py:36-40), then check that every rank ends with IDENTICAL parameters and
optimizer slots (the all-reduced gradient and the update are the same bits on every rank).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547 \
        tools/train_ranks_check.py [train.py flags]
"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from paac_b200 import train
from paac_b200.paac import PAACLearner


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    torch.distributed.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank = torch.distributed.get_rank()
    folder = tempfile.mkdtemp(prefix='paacb_ranks_%d_' % rank)
    argv = ['-g', 'synthetic', '-d', '/gpu:%d' % local, '--arch', 'NATURE', '-ec', str(16 * world), '-ew', '2',
            '--max_global_steps', str(16 * world * 5 * 4), '-df', folder + '/'] + sys.argv[1:]
    args = train.get_arg_parser().parse_args(argv)
    args.synthetic_p_terminal = 0.1
    nc, ec = train.get_network_and_environment_creator(args)
    learner = PAACLearner(nc, ec, args)
    learner.train()
    assert learner.global_step == 16 * world * 5 * 4
    for name, t in (('params', learner.network.params), ('ms', learner.engine.ms), ('mom', learner.engine.mom)):
        mine = t.clone()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        torch.distributed.all_gather(gathered, mine)
        for g in gathered:
            assert torch.equal(g, gathered[0]), 'ranks diverged in ' + name
    moved = float((learner.network.params - torch.as_tensor(nc().get_params(), device=mine.device)).abs().max().item())
    if rank == 0:
        print('train_ranks_check world=%d flags=%s: %d global steps, loss %.6f, |g| %.4f, max |dW| %.3e, ranks identical'
              % (world, ' '.join(sys.argv[1:]) or '-', learner.global_step, float(learner.last_loss.item()),
                 float(learner.last_norm.item()), moved), flush=True)
        assert moved > 0
        print('train_ranks_check ok', flush=True)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
