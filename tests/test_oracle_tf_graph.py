"""CPU: oracle/network.py + oracle/update.py against the reference's own serialized TF-1.0.1 training graph
(pretrained/breakout/checkpoints/-80000000.meta) as evaluated by oracle/tf_graph.py; outputs committed in
tests/golden/tf_graph_nips.npz.  This is what pins the model / loss / clip / RMSProp restatement."""
import json
import os

import numpy as np
import pytest

from oracle import network, update
from oracle.make_golden import gen_train_batch
from util import assert_close


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'tf_graph_nips.npz'))


def test_graph_constants(gold):
    c = json.loads(str(gold['consts']))
    assert str(gold['tf_version']) == '1.0.1'
    assert np.float32(c['local_learning/scalar']) == network.INPUT_SCALE
    assert np.float32(c['local_learning_2/Const']) == network.LOG_EPS
    assert np.float32(c['local_learning_2/Mul_4/x']) == np.float32(0.02)
    assert np.float32(c['OptimizerVariables/decay']) == np.float32(0.99)
    assert np.float32(c['OptimizerVariables/epsilon']) == np.float32(0.1)
    assert c['OptimizerVariables/momentum'] == 0.0
    assert float(gold['slot_ms_init']) == 1.0 and float(gold['slot_mom_init']) == 0.0
    ops = set(json.loads(str(gold['ops_used'])))
    assert {'Conv2D', 'Conv2DBackpropFilter', 'Conv2DBackpropInput', 'ReluGrad', 'L2Loss', 'ApplyRMSProp',
            'Softmax', 'Minimum'} <= ops


def test_two_train_steps_match_graph(gold):
    A, b = int(gold['A']), int(gold['b'])
    specs = network.param_specs('NIPS', A)
    params = network.init_params('NIPS', A, int(gold['wseed']))
    ms = {n: np.ones(s, np.float32) for n, s, _ in specs}
    mom = {n: np.zeros(s, np.float32) for n, s, _ in specs}
    lr = np.float32(gold['lr'])
    for step in range(2):
        states, acts, adv, tgt = gen_train_batch(int(gold['bseed']) + step, b, A)
        loss, grads, fwd = network.loss_and_grads(params, states, acts, adv, tgt, np.float32(0.02), 'NIPS', A)
        assert_close(fwd['pi'], gold['pi%d' % step], 1e-5, 'pi')
        assert_close(fwd['v'], gold['v%d' % step], 1e-5, 'v')
        assert abs(loss - float(gold['loss%d' % step])) <= 1e-5 * abs(float(gold['loss%d' % step]))
        glist = [grads[n] for n, _, _ in specs]
        raw_sumsq = np.asarray([np.sum(np.square(g.astype(np.float64))) for g in glist])
        assert_close(raw_sumsq, gold['raw_sumsq%d' % step], 1e-4, 'raw grad sumsq')
        head = np.concatenate([g.reshape(-1)[:64] for g in glist])
        assert_close(head, gold['raw_grad_head%d' % step], 1e-4, 'raw grad head')
        clipped, norm = update.clip_by_global_norm(glist, 3.0)
        assert abs(float(norm) - float(gold['norm%d' % step])) <= 1e-5 * float(norm)
        csq = np.asarray([np.sum(np.square(g.astype(np.float64))) for g in clipped])
        assert_close(csq, gold['clipped_sumsq%d' % step], 1e-4, 'clipped sumsq')
        for (n, _, _), g in zip(specs, clipped):
            params[n], ms[n], mom[n] = update.rmsprop_apply(params[n], ms[n], mom[n], g, lr, 0.99, 0.1)
        flat = network.flatten_params(params, 'NIPS', A)
        assert_close(flat[::997], gold['var_sample%d' % step], 1e-6, 'vars after step')
        assert_close(np.asarray([np.sum(ms[n].astype(np.float64)) for n, _, _ in specs]), gold['ms_sum%d' % step],
                     1e-6, 'rms slots')


def test_pins_for_all_shipped_games(golden_dir):
    pins = json.load(open(os.path.join(golden_dir, 'tf_graph_pins.json')))
    assert set(pins) == {'beam_rider', 'boxing', 'breakout', 'ms_pacman', 'name_this_game', 'qbert', 'seaquest',
                         'space_invaders'}
    for game, p in pins.items():
        A = p['shapes']['local_learning_2/actor_output_biases'][0]
        specs = {n: list(s) for n, s, _ in network.param_specs('NIPS', A)}
        for full, shp in p['shapes'].items():
            assert specs[full.split('/')[-1]] == shp, (game, full)
        assert np.float32(p['input_scale']) == network.INPUT_SCALE
        a = p['args']
        assert (a['initial_lr'], a['e'], a['alpha'], a['gamma'], a['clip_norm'], a['clip_norm_type'],
                a['entropy_regularisation_strength'], a['max_local_steps'], a['emulator_counts'],
                a['emulator_workers'], a['arch']) == (0.0224, 0.1, 0.99, 0.99, 3.0, 'global', 0.02, 5, 32, 8, 'NIPS')


def test_param_counts():
    # SURVEY 8a a11/a12
    assert network.param_count('NIPS', 4) == 677429 and network.param_count('NIPS', 6) == 677943
    assert network.param_count('NATURE', 4) == 1686693 and network.param_count('NATURE', 6) == 1687719
