"""train.py mirror (train.py:14-109): same flag names, dests and defaults, same
``get_network_and_environment_creator`` / ``main``.  Additions are additive flags only:
``--math`` (fp32 | tf32x3 | tf32 | bf16x3), ``--raw_frames``, ``--train_forward``, ``--graphs``, ``--summaries``,
``--synthetic_actions``, and multi-GPU launch through
torchrun (one process per GPU; RANK / WORLD_SIZE / LOCAL_RANK read from the environment).

    python -m paac_b200.train -g synthetic -d /gpu:0 --arch NATURE -ec 32 -ew 8
    python -m torch.distributed.run --nproc-per-node 8 -m paac_b200.train -g synthetic --arch NATURE -ec 256
"""
import argparse
import copy
import logging
import os
import signal
import sys

import torch

from . import environment_creator
from .paac import PAACLearner
from .policy_v_network import NaturePolicyVNetwork, NIPSPolicyVNetwork


def bool_arg(string):
    value = string.lower()
    if value == 'true':
        return True
    elif value == 'false':
        return False
    else:
        raise argparse.ArgumentTypeError("Expected True or False, but got {}".format(string))


def main(args):
    logging.debug('Configuration: {}'.format(args))

    if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not torch.distributed.is_initialized():
        local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local_rank)
        args.device = '/gpu:%d' % local_rank
        torch.distributed.init_process_group('nccl')

    network_creator, env_creator = get_network_and_environment_creator(args)

    learner = PAACLearner(network_creator, env_creator, args)

    setup_kill_signal_handler(learner)

    logging.info('Starting training')
    learner.train()
    logging.info('Finished training')


def setup_kill_signal_handler(learner):
    main_process_pid = os.getpid()

    def signal_handler(signal, frame):
        if os.getpid() == main_process_pid:
            logging.info('Signal ' + str(signal) + ' detected, cleaning up.')
            learner.cleanup()
            logging.info('Cleanup completed, shutting down...')
            sys.exit(0)

    signal.signal(signal.SIGTERM, signal_handler)
    signal.signal(signal.SIGINT, signal_handler)


def get_network_and_environment_creator(args, random_seed=3):
    args.random_seed = random_seed
    env_creator = environment_creator.EnvironmentCreator(args)
    num_actions = env_creator.num_actions
    args.num_actions = num_actions

    network_conf = {'num_actions': num_actions,
                    'entropy_regularisation_strength': args.entropy_regularisation_strength,
                    'device': args.device,
                    'clip_norm': args.clip_norm,
                    'clip_norm_type': args.clip_norm_type,
                    'math': getattr(args, 'math', 'auto'),
                    'seed': random_seed}
    if args.arch == 'NIPS':
        network = NIPSPolicyVNetwork
    else:
        network = NaturePolicyVNetwork

    def network_creator(name='local_learning'):
        nonlocal network_conf
        copied_network_conf = copy.copy(network_conf)
        copied_network_conf['name'] = name
        return network(copied_network_conf)

    return network_creator, env_creator


def get_arg_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('-g', default='pong', help='Name of game', dest='game')
    parser.add_argument('-d', '--device', default='/gpu:0', type=str, help="Device to be used ('/gpu:0', '/gpu:1',...)", dest="device")
    parser.add_argument('--rom_path', default='./atari_roms', help='Directory where the game roms are located (needed for ALE environment)', dest="rom_path")
    parser.add_argument('-v', '--visualize', default=False, type=bool_arg, help="0: no visualization of emulator; 1: all emulators, for all actors, are visualized; 2: only 1 emulator (for one of the actors) is visualized", dest="visualize")
    parser.add_argument('--e', default=0.1, type=float, help="Epsilon for the Rmsprop and Adam optimizers", dest="e")
    parser.add_argument('--alpha', default=0.99, type=float, help="Discount factor for the history/coming gradient, for the Rmsprop optimizer", dest="alpha")
    parser.add_argument('-lr', '--initial_lr', default=0.0224, type=float, help="Initial value for the learning rate. Default = 0.0224", dest="initial_lr")
    parser.add_argument('-lra', '--lr_annealing_steps', default=80000000, type=int, help="Nr. of global steps during which the learning rate will be linearly annealed towards zero", dest="lr_annealing_steps")
    parser.add_argument('--entropy', default=0.02, type=float, help="Strength of the entropy regularization term (needed for actor-critic)", dest="entropy_regularisation_strength")
    parser.add_argument('--clip_norm', default=3.0, type=float, help="If clip_norm_type is local/global, grads will be clipped at the specified maximum (avaerage) L2-norm", dest="clip_norm")
    parser.add_argument('--clip_norm_type', default="global", help="Whether to clip grads by their norm or not. Values: ignore (no clipping), local (layer-wise norm), global (global norm)", dest="clip_norm_type")
    parser.add_argument('--gamma', default=0.99, type=float, help="Discount factor", dest="gamma")
    parser.add_argument('--max_global_steps', default=80000000, type=int, help="Max. number of training steps", dest="max_global_steps")
    parser.add_argument('--max_local_steps', default=5, type=int, help="Number of steps to gain experience from before every update.", dest="max_local_steps")
    parser.add_argument('--arch', default='NIPS', help="Which network architecture to use: from the NIPS or NATURE paper", dest="arch")
    parser.add_argument('--single_life_episodes', default=False, type=bool_arg, help="If True, training episodes will be terminated when a life is lost (for games)", dest="single_life_episodes")
    parser.add_argument('-ec', '--emulator_counts', default=32, type=int, help="The amount of emulators per agent. Default is 32.", dest="emulator_counts")
    parser.add_argument('-ew', '--emulator_workers', default=8, type=int, help="The amount of emulator workers per agent. Default is 8.", dest="emulator_workers")
    parser.add_argument('-df', '--debugging_folder', default='logs/', type=str, help="Folder where to save the debugging information.", dest="debugging_folder")
    parser.add_argument('-rs', '--random_start', default=True, type=bool_arg, help="Whether or not to start with 30 noops for each env. Default True", dest="random_start")
    # ---- additions (not in the reference) ----
    parser.add_argument('--math', default='auto', choices=['auto', 'fp32', 'tf32x3', 'tf32', 'bf16x3'], help="Arithmetic of the conv/fc contractions. The reference trains in fp32; 'auto' = bf16x3: tensor cores on bf16-split operands with fp32 accumulation, within 5.4e-5 of an fp64 evaluation of the full-size gradient, 1.5e-5 of the forward (the parity bar is 1e-4). 'fp32' is the bit-faithful anchor (13x slower), 'tf32x3' the older parity-grade path, 'tf32' a speed mode outside the bar", dest="math")
    parser.add_argument('--raw_frames', default=True, type=bool_arg, help="Workers write raw frame pairs; the GPU does max-pool/resize/stack", dest="raw_frames")
    parser.add_argument('--train_forward', default='reuse', choices=['reuse', 'stepwise', 'batched'], help="Schedule of the training forward (bit-identical results): reuse the acting forwards' activations, issue it per step on a side stream, or run it inside the update like the reference", dest="train_forward")
    parser.add_argument('--graphs', default='auto', choices=['auto', 'true', 'false'], help="Replay act/update as CUDA graphs (auto: when <= 1024 environments per GPU)", dest="graphs")
    parser.add_argument('--summaries', default=False, type=bool_arg, help="Write per-update gradient statistics and episode scalars as JSONL under the debugging folder (the reference's TensorBoard summaries, opt-in)", dest="summaries")
    parser.add_argument('--summary_interval', default=100, type=int, help="Updates between gradient-statistics records", dest="summary_interval")
    parser.add_argument('--synthetic_actions', default=6, type=int, help="num_actions of the synthetic environment", dest="synthetic_actions")
    return parser


if __name__ == '__main__':
    args = get_arg_parser().parse_args()

    from . import logger_utils
    args.resolved_math = 'bf16x3' if args.math == 'auto' else args.math      # recorded: the default numerics are not fp32
    logger_utils.save_args(args, args.debugging_folder)
    logging.basicConfig(stream=sys.stdout, level=logging.DEBUG)
    logging.debug(args)

    main(args)
