"""AtariEmulator against the REFERENCE'S OWN AtariEmulator on the same game.

ALE cannot be installed here, so the game is the scripted stand-in of tests/fake_ale/ale_python_interface.py.
tests/golden/atari_fake_ale.npz was recorded by oracle/make_golden_atari.py, which imports /root/reference/atari_emulator.py
UNMODIFIED (with Pillow NEAREST for the removed scipy.misc.imresize) and drives it with the loop of emulator_runner.py:24-31:
observations, rewards and terminal flags of 280 env steps in three configurations (random no-op starts, single-life episodes,
several actor ids).  Here paac_b200's AtariEmulator plays the same actions on the same game:
  * CPU: its classic plugin API (get_initial_state / next) and its raw-frame API + the oracle's preprocessing restatement
    must reproduce the reference's observations, rewards and terminals bit for bit;
  * GPU (-m gpu): the raw frame pairs it writes, fed through paacb_preprocess_u8 with the episode-over flag as the reset
    flag (what PAACLearner.train() does), must give the same stacked observations.
"""
import os
import random
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'fake_ale'))

from oracle.make_golden_atari import CASES, Args          # the recorded configurations (no reference import at test time)

GOLD = np.load(os.path.join(HERE, 'golden', 'atari_fake_ale.npz'))


def make(case):
    from paac_b200.atari_emulator import AtariEmulator
    random.seed(case['py_seed'])
    return AtariEmulator(case['actor_id'], Args(case))


@pytest.mark.parametrize('case', CASES, ids=[c['name'] for c in CASES])
def test_classic_api_equals_the_reference_emulator(case):
    emu = make(case)
    A = len(emu.get_legal_actions())
    assert A == 4 and emu.get_noop() == [1.0, 0.0]
    obs = [emu.get_initial_state()]
    rewards, terminals = [], []
    for a in GOLD['actions_' + case['name']]:
        s, r, over = emu.next(np.eye(A)[a])
        if over:
            s = emu.get_initial_state()                       # emulator_runner.py:26-27
        obs.append(s); rewards.append(r); terminals.append(bool(over))
    assert np.array_equal(np.stack(obs), GOLD['obs_' + case['name']])
    assert np.array_equal(np.asarray(rewards, np.float64), GOLD['rewards_' + case['name']])
    assert np.array_equal(np.asarray(terminals), GOLD['terminals_' + case['name']])
    assert GOLD['terminals_a'].sum() >= 3 and GOLD['terminals_b'].sum() > GOLD['terminals_a'].sum()      # resets are exercised


def raw_rollout(case):
    """The raw-frame protocol (RawFrameEmulatorRunner._run): per step either one pair (slot 0) or, after a terminal step,
    the four pairs of the new episode's initial state; returns (first[4,2,210,160], frames[T,4,2,210,160], over[T], rewards[T])."""
    emu = make(case)
    A = len(emu.get_legal_actions())
    first = np.zeros((4, 2, 210, 160), np.uint8)
    emu.get_initial_state_raw(first)
    T = len(GOLD['actions_' + case['name']])
    frames = np.zeros((T, 4, 2, 210, 160), np.uint8)
    over = np.zeros(T, np.uint8)
    rewards = np.zeros(T, np.float64)
    for t, a in enumerate(GOLD['actions_' + case['name']]):
        r, term = emu.next_raw(np.eye(A)[a], frames[t])
        if term:
            emu.get_initial_state_raw(frames[t])
        over[t], rewards[t] = term, r
    return first, frames, over, rewards


@pytest.mark.parametrize('case', CASES, ids=[c['name'] for c in CASES])
def test_raw_protocol_plus_oracle_preprocessing_equals_the_reference_emulator(case):
    from oracle import preprocess as opre
    from paac_b200.resize_tables import ROW, COL
    first, frames, over, rewards = raw_rollout(case)
    want = GOLD['obs_' + case['name']]
    s = opre.step_states(np.zeros((1, 84, 84, 4), np.uint8), first[None], np.ones(1, np.uint8), ROW, COL)
    assert np.array_equal(s[0], want[0])
    for t in range(len(over)):
        s = opre.step_states(s, frames[t][None], over[t:t + 1], ROW, COL)
        assert np.array_equal(s[0], want[t + 1]), t
    assert np.array_equal(rewards, GOLD['rewards_' + case['name']]) and np.array_equal(over.astype(bool), GOLD['terminals_' + case['name']])


@pytest.mark.gpu
def test_raw_protocol_through_the_gpu_kernel_equals_the_reference_emulator():
    import torch
    import gpu_util as G
    net = G.make_net('NIPS', 4)
    cases = CASES[:2]
    rolls = [raw_rollout(c) for c in cases]
    T = min(len(r[2]) for r in rolls)
    n = len(cases)
    state = torch.zeros((n, 84, 84, 4), dtype=torch.uint8, device='cuda')
    state = G.preprocess(net, np.stack([r[0] for r in rolls]), 4, np.ones(n, np.uint8), state.cpu().numpy())
    assert np.array_equal(state.cpu().numpy(), np.stack([GOLD['obs_' + c['name']][0] for c in cases]))
    for t in range(T):
        state = G.preprocess(net, np.stack([r[1][t] for r in rolls]), 4, np.stack([r[2][t] for r in rolls]), state.cpu().numpy())
        assert np.array_equal(state.cpu().numpy(), np.stack([GOLD['obs_' + c['name']][t + 1] for c in cases])), t


def test_environment_creator_finds_the_emulator():
    from paac_b200.environment_creator import EnvironmentCreator
    args = Args(CASES[0])
    creator = EnvironmentCreator(args)
    assert creator.num_actions == 4
    env = creator.create_environment(0)
    assert env.supports_raw_frames and env.get_initial_state().shape == (84, 84, 4)
