"""test.py mirror (test.py:22-88): the reference's evaluation loop -- load ``<folder>/args.json`` and the newest
checkpoint under ``<folder>/checkpoints``, play ``-tc`` episodes with up to ``-np`` random no-op starts, print
mean / min / max / std of the episode returns.  Same flag names and dests; the action choice goes through
``PAACLearner.choose_next_actions`` exactly as test.py:77 does, i.e. through the B200 forward + sampling kernel.

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


def get_save_frame(name):
    import imageio                               # optional, exactly as in the reference (test.py:13-20)

    writer = imageio.get_writer(name + '.gif', fps=30)

    def get_frame(frame):
        writer.append_data(frame)

    return get_frame


def get_arg_parser():
    parser = argparse.ArgumentParser()
    parser.add_argument('-f', '--folder', type=str, help="Folder where to save the debugging information.", dest="folder", required=True)
    parser.add_argument('-tc', '--test_count', default='1', type=int, help="The amount of tests to run on the given network", dest="test_count")
    parser.add_argument('-np', '--noops', default=30, type=int, help="Maximum amount of no-ops to use", dest="noops")
    parser.add_argument('-gn', '--gif_name', default=None, type=str, help="If provided, a gif will be produced and stored with this name", dest="gif_name")
    parser.add_argument('-gf', '--gif_folder', default='', type=str, help="The folder where to save gifs.", dest="gif_folder")
    parser.add_argument('-d', '--device', default='/gpu:0', type=str, help="Device to be used ('/gpu:0', '/gpu:1',...)", dest="device")
    return parser


def evaluate(args):
    """-> float32 array of the test_count episode returns (test.py:32-82)."""
    arg_file = os.path.join(args.folder, 'args.json')
    device = args.device
    for k, v in logger_utils.load_args(arg_file).items():
        setattr(args, k, v)
    args.max_global_steps = 0
    df = args.folder
    args.debugging_folder = '/tmp/logs'
    args.device = device

    args.random_start = False
    args.single_life_episodes = False
    if args.gif_name:
        args.visualize = 1

    args.actor_id = 0
    rng = np.random.RandomState(int(time.time()))
    args.random_seed = rng.randint(1000)

    network_creator, env_creator = get_network_and_environment_creator(args)
    network = network_creator()

    def set_state(state):
        for n, t in network.variables().items():
            t.copy_(state[n])
        network.params_changed()

    saver = Saver(lambda: {n: t.detach().cpu() for n, t in network.variables().items()}, set_state,
                  scope=getattr(network, 'name', 'local_learning'))

    environments = [env_creator.create_environment(i) for i in range(args.test_count)]
    if args.gif_name:
        for i, environment in enumerate(environments):
            environment.on_new_frame = get_save_frame(os.path.join(args.gif_folder, args.gif_name + str(i)))

    sess = Session()
    network.init(os.path.join(df, 'checkpoints'), saver, sess)
    states = np.asarray([environment.get_initial_state() for environment in environments])
    if args.noops != 0:
        for i, environment in enumerate(environments):
            for _ in range(random.randint(0, args.noops)):
                state, _, _ = environment.next(environment.get_noop())
                states[i] = state

    episodes_over = np.zeros(args.test_count, dtype=bool)
    rewards = np.zeros(args.test_count, dtype=np.float32)
    while not all(episodes_over):
        actions, _, _ = PAACLearner.choose_next_actions(network, env_creator.num_actions, states, sess)
        for j, environment in enumerate(environments):
            if episodes_over[j]:
                continue                          # the reference keeps stepping finished episodes; their returns are final here
            state, r, episode_over = environment.next(actions[j])
            states[j] = state
            rewards[j] += r
            episodes_over[j] = episode_over
    sess.close()
    return rewards


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
