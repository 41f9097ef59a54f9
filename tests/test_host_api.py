"""CPU: the host-side mirror of the reference plugin API (names, argument order, protocol, defaults)."""
import argparse
import os
import re

import numpy as np
import pytest

from oracle import preprocess as opre
from oracle.make_golden import ScriptedEnv, scripted_actions
from paac_b200 import environment, emulator_runner, runners, synthetic_emulator
from paac_b200.resize_tables import ROW, COL


def test_base_environment_contract():
    e = environment.BaseEnvironment()
    for call in (e.get_initial_state, lambda: e.next([1, 0]), e.get_legal_actions, e.get_noop):
        with pytest.raises(NotImplementedError):
            call()
    assert e.on_new_frame(None) is None


def test_runners_protocol_matches_reference_golden(golden_dir):
    """Same scripted envs, same actions, through OUR Runners/EmulatorRunner: identical shared-buffer
    contents to what the reference's runners.py / emulator_runner.py produced (runner_golden.npz)."""
    g = np.load(os.path.join(golden_dir, 'runner_golden.npz'))
    n, W, A, period, steps = (int(g[k]) for k in ('n', 'W', 'A', 'period', 'steps'))
    emus = np.asarray([ScriptedEnv(i, A, period) for i in range(n)])
    variables = [np.asarray([e.get_initial_state() for e in emus], dtype=np.uint8), np.zeros(n, dtype=np.float32),
                 np.asarray([False] * n, dtype=np.float32), np.zeros((n, A), dtype=np.float32)]
    rr = runners.Runners(emulator_runner.EmulatorRunner, emus, W, variables)
    sh_states, sh_rew, sh_over, sh_act = rr.get_shared_variables()
    assert sh_states.dtype == np.uint8            # App. E.1: true uint8 (the reference shares uint32)
    assert str(g['shared_state_dtype']) == 'uint32'
    rr.start()
    try:
        acts = scripted_actions(int(g['act_seed']), steps, n, A)
        assert (sh_states == g['states'][0]).all()
        for t in range(steps):
            sh_act[:] = np.eye(A, dtype=np.float32)[acts[t]]
            rr.update_environments()
            rr.wait_updated()
            assert (sh_states == g['states'][t + 1]).all()
            assert (sh_rew == g['rewards'][t]).all() and (sh_over == g['over'][t]).all()
    finally:
        rr.stop()
        for r in rr.runners:
            r.join(5)
    assert g['over'].sum() > 0                    # the fixture crosses episode boundaries


class _Args(object):
    random_seed = 3
    synthetic_actions = 6
    synthetic_p_terminal = 0.2


def test_raw_frame_protocol_equals_classic_protocol():
    """A worker on the raw-frame protocol + the oracle pipeline == the same env on the classic protocol."""
    a = synthetic_emulator.SyntheticEmulator(5, _Args())
    b = synthetic_emulator.SyntheticEmulator(5, _Args())
    s_classic = a.get_initial_state()
    slots = np.zeros((1, 4, 2, 210, 160), np.uint8)
    b.get_initial_state_raw(slots[0])
    s_raw = opre.step_states(np.zeros((1, 84, 84, 4), np.uint8), slots, np.ones(1, np.uint8), ROW, COL)
    assert (s_raw[0] == s_classic).all()
    rng = np.random.RandomState(0)
    n_term = 0
    for t in range(40):
        act = np.eye(6)[rng.randint(6)]
        s1, r1, d1 = a.next(act)
        r2, d2 = b.next_raw(act, slots[0])
        assert r1 == r2 and d1 == d2
        if d1:
            s1 = a.get_initial_state()
            b.get_initial_state_raw(slots[0])
            n_term += 1
        s_raw = opre.step_states(s_raw, slots, np.asarray([d2], np.uint8), ROW, COL)
        assert (s_raw[0] == s1).all(), t
    assert n_term > 0


def test_arg_parser_defaults_match_reference():
    from paac_b200 import train
    d = vars(train.get_arg_parser().parse_args([]))
    ref = dict(game='pong', device='/gpu:0', rom_path='./atari_roms', visualize=False, e=0.1, alpha=0.99,
               initial_lr=0.0224, lr_annealing_steps=80000000, entropy_regularisation_strength=0.02, clip_norm=3.0,
               clip_norm_type='global', gamma=0.99, max_global_steps=80000000, max_local_steps=5, arch='NIPS',
               single_life_episodes=False, emulator_counts=32, emulator_workers=8, debugging_folder='logs/',
               random_start=True)
    for k, v in ref.items():
        assert d[k] == v, k
    assert train.bool_arg('True') is True and train.bool_arg('false') is False
    with pytest.raises(argparse.ArgumentTypeError):
        train.bool_arg('maybe')


def test_device_string_rules():
    from paac_b200 import networks, _lib
    assert networks.parse_device('/gpu:3').index == 3
    with pytest.raises(_lib.PaacbError):
        networks.parse_device('/cpu:0')           # the product has no CPU path


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'paac_b200')
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f


def test_saver_roundtrip(tmp_path):
    import torch
    from paac_b200.session import Saver
    state = {'w': torch.arange(6.0)}
    holder = {}
    sv = Saver(lambda: state, holder.update, max_to_keep=1)
    sv.save(None, str(tmp_path) + '/', 10)
    sv.save(None, str(tmp_path) + '/', 1000000)
    path = Saver.latest_checkpoint(str(tmp_path) + '/')
    # a TensorFlow-1 bundle prefix (tf.train.latest_checkpoint returns the prefix too): .index + .data + `checkpoint`
    assert path.endswith('-1000000') and sorted(os.listdir(tmp_path)) == ['-1000000.data-00000-of-00001', '-1000000.index',
                                                                           'checkpoint']
    sv.restore(None, path)
    # checkpoints written by earlier versions of this package (torch.save) still restore
    torch.save({'w': torch.arange(6.0) + 1}, str(tmp_path / '-2000000.pt'))
    os.remove(str(tmp_path / 'checkpoint'))
    legacy = Saver.latest_checkpoint(str(tmp_path) + '/')
    assert legacy.endswith('-2000000.pt')
    sv.restore(None, legacy)
    assert torch.equal(holder['w'], torch.arange(6.0) + 1)
    sv.restore(None, path)
    assert int(path[path.rindex('-') + 1:].split('.')[0]) == 1000000 and torch.equal(holder['w'], state['w'])
