"""Evaluation of a trained run: stands in for the reference's test.py (same flags ``-f -tc -np -gn -gf -d``, same printed
summary).  It loads ``<folder>/args.json`` and the newest TensorFlow-1 bundle under ``<folder>/checkpoints`` (written by
this implementation or by the reference, session.Saver), plays ``-tc`` episodes in parallel with up to ``-np`` random no-op
starts and reports mean / min / max / std of the episode returns.  Every action comes from
``PAACLearner.choose_next_actions`` (test.py:77), i.e. from the B200 forward + sampling kernels.

    python -m paac_b200.test -f logs/ -tc 5
"""
import argparse
import os
import random
import time

import numpy as np

from . import logger_utils
from .paac import PAACLearner
from .session import Saver, Session
from .train import get_network_and_environment_creator


def get_arg_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('-f', '--folder', type=str, help="Folder where to save the debugging information.", dest="folder", required=True)
    parser.add_argument('-tc', '--test_count', default='1', type=int, help="The amount of tests to run on the given network", dest="test_count")
    parser.add_argument('-np', '--noops', default=30, type=int, help="Maximum amount of no-ops to use", dest="noops")
    parser.add_argument('-gn', '--gif_name', default=None, type=str, help="If provided, a gif will be produced and stored with this name", dest="gif_name")
    parser.add_argument('-gf', '--gif_folder', default='', type=str, help="The folder where to save gifs.", dest="gif_folder")
    parser.add_argument('-d', '--device', default='/gpu:0', type=str, help="Device to be used ('/gpu:0', '/gpu:1',...)", dest="device")
    return parser


def run_config(args):
    """The training run's args.json with the evaluation overrides of test.py:33-52 applied on top."""
    device, folder = args.device, args.folder
    for key, value in logger_utils.load_args(os.path.join(folder, 'args.json')).items():
        setattr(args, key, value)
    args.device, args.folder = device, folder
    args.max_global_steps = 0
    args.debugging_folder = '/tmp/logs'
    args.random_start = False                  # the no-op starts below replace ALE's own
    args.single_life_episodes = False
    args.visualize = 1 if args.gif_name else getattr(args, 'visualize', False)
    args.actor_id = 0
    args.random_seed = int(np.random.RandomState(int(time.time())).randint(1000))
    return args


def restored_network(args):
    network_creator, env_creator = get_network_and_environment_creator(args)
    network = network_creator()

    def write(state):
        for name, tensor in network.variables().items():
            tensor.copy_(state[name])
        network.params_changed()

    saver = Saver(lambda: {n: t.detach().cpu() for n, t in network.variables().items()}, write,
                  scope=getattr(network, 'name', 'local_learning'))
    network.init(os.path.join(args.folder, 'checkpoints'), saver, None)
    return network, env_creator


def gif_sink(path):
    import imageio                               # optional dependency, only with -gn
    writer = imageio.get_writer(path + '.gif', fps=30)
    return writer.append_data


def play(network, environments, num_actions, max_noops):
    """All episodes advance in lock-step until the last one ends; a finished episode's return is final (the reference keeps
    stepping finished environments and adds their rewards, test.py:78-82 -- a quirk, not reproduced)."""
    session = Session()
    states = np.asarray([env.get_initial_state() for env in environments])
    for i, env in enumerate(environments):
        for _ in range(random.randint(0, max_noops) if max_noops else 0):
            states[i] = env.next(env.get_noop())[0]
    returns = np.zeros(len(environments), dtype=np.float32)
    live = np.ones(len(environments), dtype=bool)
    while live.any():
        actions, _, _ = PAACLearner.choose_next_actions(network, num_actions, states, session)
        for j in np.nonzero(live)[0]:
            states[j], reward, over = environments[j].next(actions[j])
            returns[j] += reward
            live[j] = not over
    session.close()
    return returns


def evaluate(args):
    """-> float32 array of the test_count episode returns."""
    args = run_config(args)
    network, env_creator = restored_network(args)
    environments = [env_creator.create_environment(i) for i in range(args.test_count)]
    if args.gif_name:
        for i, env in enumerate(environments):
            env.on_new_frame = gif_sink(os.path.join(args.gif_folder, args.gif_name + str(i)))
    return play(network, environments, env_creator.num_actions, args.noops)


def main(argv=None):
    args = get_arg_parser().parse_args(argv)
    rewards = evaluate(args)
    print('Performed {} tests for {}.'.format(args.test_count, args.game))
    print('Mean: {0:.2f}'.format(np.mean(rewards)))
    print('Min: {0:.2f}'.format(np.min(rewards)))
    print('Max: {0:.2f}'.format(np.max(rewards)))
    print('Std: {0:.2f}'.format(np.std(rewards)))
    return rewards


if __name__ == '__main__':
    main()
