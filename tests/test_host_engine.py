"""Host logic of RolloutEngine's training-forward schedules on CPU, with a recording stand-in for the network and
the C-ABI library (no compute calls): which forwards are issued, on which samples, into which workspace slice."""
import pytest
import torch

from paac_b200.engine import RolloutEngine


class FakeLib(object):
    def paacb_backward_workspace_floats(self, ctx, b):
        return 8 * b

    def paacb_optimizer_workspace_floats(self, ctx):
        return 16

    def paacb_grad_tail_offset(self, ctx):
        return 10


class FakeNet(object):
    def __init__(self):
        self._lib, self.ctx = FakeLib(), None
        self.torch_device = torch.device('cpu')
        self.num_actions, self.entropy_regularisation_strength, self.param_count = 4, 0.02, 40
        self.calls = []

    def workspace_floats(self, b):
        return 8 * b

    def forward(self, states, pi, v, ws, uniforms=None, actions=None, onehot=None, ws_capacity=None, ws_first=0, rng=None,
                draw=0, first_sample=0):
        self.calls.append(dict(b=states.shape[0], cap=ws_capacity, first=ws_first, sample=uniforms is not None or rng is not None,
                               draw=draw, first_sample=first_sample,
                               ws=ws.data_ptr(), pi=pi.data_ptr(), v=v.data_ptr()))


def make(mode, N=12, T=3):
    net = FakeNet()
    return net, RolloutEngine(net, N, T, train_forward=mode)


def test_rejects_unknown_schedule():
    with pytest.raises(ValueError):
        make('eager')


def test_batched_acts_in_the_acting_workspace():
    net, eng = make('batched')
    eng.act(1)
    eng.act(1, 0, 4)
    assert [c['cap'] for c in net.calls] == [None, None] and net.calls[0]['ws'] == eng.act_ws.data_ptr()
    assert net.calls[1]['ws'] != eng.fwd_ws.data_ptr() and net.calls[1]['b'] == 4
    with pytest.raises(RuntimeError):
        eng.train_forward_step(0)


def test_reuse_writes_the_training_workspace_and_aliases_values():
    net, eng = make('reuse')
    eng.act(2, 4, 12)
    c = net.calls[0]
    assert (c['b'], c['cap'], c['first'], c['sample']) == (8, eng.B, 2 * eng.N + 4, True)
    assert c['ws'] == eng.fwd_ws.data_ptr()
    assert c['pi'] == eng.pi[2 * eng.N + 4:].data_ptr() and c['v'] == eng.v[2 * eng.N + 4:].data_ptr()
    assert eng.values.data_ptr() == eng.v.data_ptr() and tuple(eng.values.shape) == (eng.T, eng.N)
    eng.set_train_forward('batched')
    assert eng.values.data_ptr() != eng.v.data_ptr()


def test_stepwise_fills_exactly_the_missing_slices():
    net, eng = make('stepwise')
    eng.train_forward_step(0)                 # whole step 0
    eng.train_forward_step(2, 3, 7)           # a middle slice of step 2; step 1 not issued at all
    eng.train_forward_step(2, 9, 12)
    net.calls.clear()
    eng._finish_stepwise()
    got = sorted((c['first'], c['b']) for c in net.calls)
    N = eng.N
    assert got == [(1 * N, N), (2 * N, 3), (2 * N + 7, 2)]
    assert all(c['cap'] == eng.B and c['ws'] == eng.fwd_ws.data_ptr() and not c['sample'] for c in net.calls)
    # the bookkeeping is per update
    net.calls.clear()
    eng._finish_stepwise()
    assert sorted((c['first'], c['b']) for c in net.calls) == [(t * N, N) for t in range(eng.T)]


def test_bootstrap_slices_are_completed_by_update_bookkeeping():
    net, eng = make('batched')
    eng.bootstrap(4, 8)
    net.calls.clear()
    # forward_backward's bootstrap bookkeeping (the C calls that follow need a GPU)
    assert eng._missing_bootstraps() == [(0, 4), (8, eng.N)]
    eng.bootstrap(0, 4); eng.bootstrap(8, eng.N)
    assert [(c['b']) for c in net.calls] == [4, eng.N - 8]
    assert all(not c['sample'] and c['cap'] is None for c in net.calls)
