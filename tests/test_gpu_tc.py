"""GPU: the tcgen05 tensor-core paths against the oracle.  PAACB_MATH_BF16X3 (bf16-split operands, TMA-fed
patch-resident implicit GEMMs; Nature architecture) and PAACB_MATH_TF32X3 are parity modes and meet the 1e-4 bar
(scaled max error, see util.py); plain TF32 rounds operands to 11 bits and is held to 3e-3 -- a speed mode only."""
import numpy as np
import pytest
import torch

from oracle import network, update
from paac_b200.engine import RolloutEngine
from util import assert_close, rel_err
import gpu_util as G

pytestmark = pytest.mark.gpu

TOL = {'bf16x3': 1e-4, 'tf32x3': 1e-4, 'tf32': 3e-3}
MODES = ['bf16x3', 'tf32x3', 'tf32']


def skip_unsupported(math, arch):
    if math == 'bf16x3' and arch == 'NIPS':
        pytest.skip('PAACB_MATH_BF16X3 covers the Nature architecture (see test_bf16x3_rejects_nips)')


@pytest.mark.parametrize('math', MODES)
@pytest.mark.parametrize('arch,A,b', [('NATURE', 6, 1), ('NATURE', 6, 130), ('NIPS', 4, 33), ('NATURE', 18, 517)])
def test_forward_tc_vs_oracle(math, arch, A, b):
    skip_unsupported(math, arch)
    net = G.make_net(arch, A, seed=3, math=math)
    params = network.unflatten_params(net.get_params(), arch, A)
    rng = np.random.RandomState(b)
    states = rng.randint(0, 256, (b, 84, 84, 4)).astype(np.uint8)
    out = G.forward(net, states)
    ref = network.forward(params, states, arch, dtype=torch.float64, keep=True)
    acts = G.layer_acts(net, out['ws'], b)
    for i, a in enumerate(ref['acts']):
        assert_close(acts[i], a.numpy(), TOL[math], 'conv%d activation' % (i + 1))
    assert_close(acts[-1], ref['h'].numpy(), TOL[math], 'hidden fc')
    assert_close(out['pi'].cpu().numpy(), ref['pi'].numpy(), TOL[math], 'pi')
    assert_close(out['v'].cpu().numpy(), ref['v'].numpy(), TOL[math], 'v')


@pytest.mark.parametrize('math', MODES)
@pytest.mark.parametrize('arch,A,b', [('NATURE', 6, 5), ('NATURE', 6, 160), ('NIPS', 4, 97), ('NATURE', 4, 1111)])
def test_backward_tc_vs_autograd(math, arch, A, b):
    """Gradients (and every layer's dZ) against fp64 autograd whose ReLU masks are the GPU's own activations
    (oracle.network.masked_loss_and_grads): like-for-like, immune to a pre-activation rounding across zero."""
    skip_unsupported(math, arch)
    net = G.make_net(arch, A, seed=11, math=math)
    params = network.unflatten_params(net.get_params(), arch, A)
    rng = np.random.RandomState(b + A)
    states = rng.randint(0, 256, (b, 84, 84, 4)).astype(np.uint8)
    acts = rng.randint(0, A, b)
    adv = rng.randn(b).astype(np.float32); tgt = rng.randn(b).astype(np.float32)
    fwd = G.forward(net, states)
    masks = [x > 0 for x in G.layer_acts(net, fwd['ws'], b)]
    g64, dzs, f64 = network.masked_loss_and_grads(params, states, acts, adv, tgt, 0.02, arch, A, masks)
    _, dz, dv = network.closed_form_head_grads(f64['logits'], f64['v'], acts, adv, tgt, np.float32(0.02))
    flat, bws = G.backward(net, fwd, dz, dv)
    got = network.unflatten_params(flat, arch, A)
    for name, _, _ in network.param_specs(arch, A):
        assert_close(got[name], g64[name], TOL[math], name)
    for i, (d, got_dz) in enumerate(zip(dzs, G.layer_acts(net, bws, b))):
        assert_close(got_dz.reshape(d.shape), d, TOL[math], 'dZ of layer %d' % i)


def test_bf16x3_rejects_nips():
    from paac_b200 import _lib
    with pytest.raises(_lib.PaacbError):
        G.make_net('NIPS', 4, seed=3, math='bf16x3')


@pytest.mark.parametrize('math', ['bf16x3', 'tf32x3'])
def test_engine_update_tc_vs_oracle_composite(math):
    arch, A, N, T = 'NATURE', 6, 16, 5
    net = G.make_net(arch, A, seed=5, math=math)
    eng = RolloutEngine(net, N, T, seed=9)
    rng = np.random.RandomState(0)
    states = rng.randint(0, 256, (T + 1, N, 84, 84, 4)).astype(np.uint8)
    eng.states.copy_(G.dev(states))
    eng.draw_uniforms()
    for t in range(T):
        eng.act(t)
    rewards = rng.choice([-2.0, 0.0, 1.0], size=(T, N)).astype(np.float32)
    over = (rng.random_sample((T, N)) < 0.15).astype(np.float32)
    eng.rewards.copy_(G.dev(rewards)); eng.over.copy_(G.dev(over))
    p0 = net.get_params()
    eng.update(0.0224)
    torch.cuda.synchronize()
    params = network.unflatten_params(p0, arch, A)
    B = T * N
    acts = eng.actions.cpu().numpy().reshape(-1)
    y, adv = update.nstep_returns(rewards, over, eng.values.cpu().numpy(), eng.boot_v.cpu().numpy(), 0.99)
    loss, _, _ = network.loss_and_grads(params, states[:T].reshape(B, 84, 84, 4), acts, adv.reshape(-1), y.reshape(-1),
                                        np.float32(0.02), arch, A)
    assert abs(eng.loss.item() - loss) <= 1e-4 * max(1, abs(loss))
    masks = [x > 0 for x in G.layer_acts(net, eng.fwd_ws, B)]
    grads, _, _ = network.masked_loss_and_grads(params, states[:T].reshape(B, 84, 84, 4), acts, adv.reshape(-1), y.reshape(-1),
                                                np.float32(0.02), arch, A, masks)
    specs = network.param_specs(arch, A)
    clipped, norm = update.clip_by_global_norm([grads[n] for n, _, _ in specs], 3.0)
    assert abs(eng.norm.item() - float(norm)) <= 1e-4 * float(norm)
    new = {}
    for (n, s, _), gc in zip(specs, clipped):
        new[n], _, _ = update.rmsprop_apply(params[n], np.ones(s, np.float32), np.zeros(s, np.float32), gc, 0.0224, 0.99, 0.1)
    want = network.flatten_params(new, arch, A)
    got = net.get_params()
    assert_close(got, want, 1e-5, 'post-RMSProp weights')
    assert_close(got - p0, want - p0, 2e-3, 'weight delta')


def test_sliced_act_observe_equals_whole_batch():
    """RolloutEngine.act / observe_frames on contiguous environment slices (what the pipelined end-to-end loop issues on
    two streams) produce the same actions, values and states as the whole-batch calls."""
    arch, A, N, T = 'NATURE', 6, 24, 2
    rng = np.random.RandomState(7)
    s0 = rng.randint(0, 256, (N, 84, 84, 4)).astype(np.uint8)
    frames = G.dev(rng.randint(0, 256, (T, N, 1, 2, 210, 160)).astype(np.uint8))
    rew = G.dev(rng.choice([-1.0, 0.0, 1.0], size=(T, N)).astype(np.float32))
    over = G.dev(np.zeros((T, N), np.float32))
    outs = []
    for slices in ([(0, N)], [(0, 8), (8, 24)]):
        net = G.make_net(arch, A, seed=5, math='bf16x3')
        eng = RolloutEngine(net, N, T, seed=9)
        eng.states[0].copy_(G.dev(s0))
        eng.draw_uniforms()
        for t in range(T):
            for lo, hi in slices:
                eng.act(t, lo, hi)
            for lo, hi in slices:
                eng.observe_frames(t, frames[t, lo].data_ptr(), 1, None, rew[t], over[t], lo, hi)
        torch.cuda.synchronize()
        outs.append((eng.actions.cpu().numpy(), eng.values.cpu().numpy(), eng.states.cpu().numpy(), eng.rewards.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][2], outs[1][2])
    assert np.array_equal(outs[0][3], outs[1][3])
    assert_close(outs[1][1], outs[0][1], 1e-6, 'values of sliced vs whole-batch forward')
