"""Oracle (TEST INFRASTRUCTURE ONLY): tests/golden/atari_fake_ale.npz by RUNNING THE REFERENCE'S OWN AtariEmulator.

    python -m oracle.make_golden_atari          (build container only: needs /root/reference)

/root/reference/atari_emulator.py is imported UNMODIFIED.  Its two absent dependencies are stood in for by test
infrastructure: ``ale_python_interface`` by the scripted game of tests/fake_ale/ (ALE itself cannot be installed here), and
``scipy.misc.imresize(img, (84, 84), interp='nearest')`` -- removed from SciPy -- by Pillow's Image.resize(NEAREST), which is
what imresize called.  Everything the reference does around the emulator is therefore the reference's own
code: action repeat with the last two frames pooled (atari_emulator.py:77-86), np.amax + resize (:69-75), ObservationPool
(environment.py:58-75), reset with random no-ops and four initial repeats (:60-67, 88-96), terminal / lost-life rule
(:108-115), and the reset-on-terminal rule of emulator_runner.py:26-27 which the driver loop below applies.
"""
import os
import random
import sys
import types

import numpy as np

REF = '/root/reference'
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')

CASES = [dict(name='a', random_seed=3, actor_id=0, single_life_episodes=False, random_start=True, steps=150, py_seed=11, act_seed=5),
         dict(name='b', random_seed=3, actor_id=2, single_life_episodes=True, random_start=True, steps=90, py_seed=12, act_seed=6),
         dict(name='c', random_seed=7, actor_id=1, single_life_episodes=False, random_start=False, steps=40, py_seed=13, act_seed=7)]


class Args(object):
    def __init__(self, case):
        self.random_seed = case['random_seed']
        self.rom_path = './atari_roms'
        self.game = 'fake'
        self.random_start = case['random_start']
        self.single_life_episodes = case['single_life_episodes']
        self.visualize = False


def drive(emulator_cls, case):
    """The loop of emulator_runner.py:24-31 for one emulator: actions from a seeded stream, a fresh initial state on terminal."""
    random.seed(case['py_seed'])                           # the reference draws its no-op count from Python's global stream
    emu = emulator_cls(case['actor_id'], Args(case))
    A = len(emu.get_legal_actions())
    rng = np.random.RandomState(case['act_seed'])
    obs = [emu.get_initial_state()]
    rewards, terminals, actions = [], [], []
    for _ in range(case['steps']):
        a = int(rng.randint(0, A))
        new_s, reward, over = emu.next(np.eye(A)[a])
        if over:
            new_s = emu.get_initial_state()
        obs.append(new_s); rewards.append(reward); terminals.append(bool(over)); actions.append(a)
    return np.stack(obs).astype(np.uint8), np.asarray(rewards, np.float64), np.asarray(terminals), np.asarray(actions, np.int32)


def main():
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'fake_ale'))
    sys.path.insert(0, REF)
    from PIL import Image

    def imresize(img, size, interp='nearest'):
        assert interp == 'nearest' and img.dtype == np.uint8
        return np.asarray(Image.fromarray(img).resize((size[1], size[0]), Image.NEAREST))
    import scipy
    misc = types.ModuleType('scipy.misc')
    misc.imresize = imresize
    sys.modules['scipy.misc'] = misc
    scipy.misc = misc
    import atari_emulator as ref                           # the reference module, unmodified
    assert ref.__file__.startswith(REF)
    out = {}
    for case in CASES:
        obs, rew, term, act = drive(ref.AtariEmulator, case)
        out['obs_' + case['name']] = obs
        out['rewards_' + case['name']] = rew
        out['terminals_' + case['name']] = term
        out['actions_' + case['name']] = act
        print(case['name'], obs.shape, 'episodes ended:', int(term.sum()), 'reward sum', rew.sum(), 'nonzero pixels', int((obs > 0).sum()))
    np.savez_compressed(os.path.join(OUT, 'atari_fake_ale.npz'), **out)
    print('wrote', os.path.join(OUT, 'atari_fake_ale.npz'), os.path.getsize(os.path.join(OUT, 'atari_fake_ale.npz')), 'bytes')


if __name__ == '__main__':
    main()
