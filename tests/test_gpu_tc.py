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
    """every tensor-core mode covers both architectures since round 2 (NIPS: 64-byte units, SWIZZLE_64B variants)"""


@pytest.mark.parametrize('math', MODES)
@pytest.mark.parametrize('arch,A,b', [('NATURE', 6, 1), ('NATURE', 6, 130), ('NIPS', 4, 33), ('NIPS', 6, 300), ('NATURE', 18, 517)])
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
@pytest.mark.parametrize('arch,A,b', [('NATURE', 6, 5), ('NATURE', 6, 160), ('NIPS', 4, 97), ('NIPS', 6, 160), ('NATURE', 4, 1111)])
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


def test_auto_math_is_bf16x3_for_both_architectures():
    for arch in ('NIPS', 'NATURE'):
        net = G.make_net(arch, 4, seed=3, math='auto')
        assert net.math == 'bf16x3' and net._lib.paacb_get_math(net.ctx) == 3


@pytest.mark.parametrize('arch,math', [('NATURE', 'bf16x3'), ('NATURE', 'tf32x3'), ('NIPS', 'bf16x3')])
def test_engine_update_tc_vs_oracle_composite(arch, math):
    A, N, T = 6, 16, 5
    net = G.make_net(arch, A, seed=5, math=math)
    eng = RolloutEngine(net, N, T, seed=9)
    rng = np.random.RandomState(0)
    states = rng.randint(0, 256, (T + 1, N, 84, 84, 4)).astype(np.uint8)
    eng.set_states(G.dev(states))
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
        eng.state(0).copy_(G.dev(s0))
        eng.draw_uniforms()
        for t in range(T):
            for lo, hi in slices:
                eng.act(t, lo, hi)
            for lo, hi in slices:
                eng.observe_frames(t, frames[t, lo].data_ptr(), 1, None, rew[t], over[t], lo, hi)
        torch.cuda.synchronize()
        outs.append((eng.actions.cpu().numpy(), eng.values.cpu().numpy(), eng.get_states().cpu().numpy(), eng.rewards.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][2], outs[1][2])
    assert np.array_equal(outs[0][3], outs[1][3])
    assert_close(outs[1][1], outs[0][1], 1e-6, 'values of sliced vs whole-batch forward')


@pytest.mark.parametrize('math', ['bf16x3', 'tf32x3', 'fp32'])
def test_training_forward_schedules_are_bit_identical(math):
    """The training forward of PAAC is the concatenation of the acting forwards under unchanged parameters (paac.py:92,112,151).
    'batched' (one forward over T*N samples in update()), 'stepwise' (paacb_policy_forward_at per step / slice) and 'reuse'
    (the acting forward writes the training workspace) must leave the SAME BITS in the activation workspace, pi and v, and
    the same update (gradient sums use atomics, so the post-update weights are compared to rounding)."""
    arch, A, N, T = 'NATURE', 6, 40, 3
    rng = np.random.RandomState(11)
    states = rng.randint(0, 256, (T + 1, N, 84, 84, 4)).astype(np.uint8)
    rewards = rng.choice([-1.0, 0.0, 1.0], size=(T, N)).astype(np.float32)
    over = (rng.random_sample((T, N)) < 0.2).astype(np.float32)
    res = {}
    for mode in ('batched', 'stepwise', 'reuse'):
        net = G.make_net(arch, A, seed=5, math=math)
        eng = RolloutEngine(net, N, T, seed=9, train_forward=mode)
        eng.set_states(G.dev(states))
        eng.draw_uniforms()
        for t in range(T):
            eng.act(t, 0, 16)
            eng.act(t, 16, N)
            if mode == 'stepwise' and t != 1:                # step 1 is left to update(); step 2 is issued in two slices
                if t == 2:
                    eng.train_forward_step(t, 0, 24)
                    eng.train_forward_step(t, 24, N)
                else:
                    eng.train_forward_step(t)
        eng.rewards.copy_(G.dev(rewards)); eng.over.copy_(G.dev(over))
        eng.update(0.0224)
        torch.cuda.synchronize()
        res[mode] = dict(ws=eng.fwd_ws.view(torch.int32).cpu().numpy().copy(), pi=eng.pi.cpu().numpy().copy(),
                         v=eng.v.cpu().numpy().copy(), values=eng.values.cpu().numpy().copy(),
                         actions=eng.actions.cpu().numpy().copy(), y=eng.y.cpu().numpy().copy(),
                         loss=eng.loss.item(), params=net.get_params())
    ref = res['batched']
    for mode in ('stepwise', 'reuse'):
        r = res[mode]
        assert np.array_equal(r['actions'], ref['actions']), mode
        assert np.array_equal(r['ws'], ref['ws']), mode + ': activation workspace differs'
        assert np.array_equal(r['pi'].view(np.int32), ref['pi'].view(np.int32)), mode
        assert np.array_equal(r['v'].view(np.int32), ref['v'].view(np.int32)), mode
        assert np.array_equal(r['values'].view(np.int32), ref['values'].view(np.int32)), mode
        assert np.array_equal(r['y'].view(np.int32), ref['y'].view(np.int32)), mode
        assert abs(r['loss'] - ref['loss']) <= 1e-6 * max(1.0, abs(ref['loss']))
        assert_close(r['params'], ref['params'], 1e-6, mode + ': post-update weights')


def test_cached_weight_images_follow_the_parameters():
    """The bf16x3 forward reuses the operand images of the weights between calls (paacb_params_changed contract):
    after set_params() and after paacb_clip_rmsprop the next forward must see the NEW parameters."""
    arch, A, N, T = 'NATURE', 6, 8, 2
    rng = np.random.RandomState(3)
    states = G.dev(rng.randint(0, 256, (N, 84, 84, 4)).astype(np.uint8))

    def fwd(net):
        pi = torch.empty((N, A), device='cuda'); v = torch.empty((N,), device='cuda')
        ws = torch.empty((net.workspace_floats(N),), device='cuda')
        net.forward(states, pi, v, ws)
        torch.cuda.synchronize()
        return pi.cpu().numpy(), v.cpu().numpy()

    net = G.make_net(arch, A, seed=5, math='bf16x3')
    other = G.make_net(arch, A, seed=6, math='bf16x3')
    pi0, v0 = fwd(net)
    net.set_params(other.get_params())                      # host write -> params_changed()
    pi1, v1 = fwd(net)
    pi_o, v_o = fwd(other)
    assert np.array_equal(pi1, pi_o) and np.array_equal(v1, v_o)
    assert not np.array_equal(v0, v1)
    # an update through the library refreshes the images in-stream
    eng = RolloutEngine(net, N, T, seed=9)
    eng.set_states(G.dev(rng.randint(0, 256, (T + 1, N, 84, 84, 4)).astype(np.uint8)))
    eng.draw_uniforms()
    for t in range(T):
        eng.act(t)
    eng.rewards.fill_(1.0)
    eng.update(0.0224)
    pi2, v2 = fwd(net)
    fresh = G.make_net(arch, A, seed=1, math='bf16x3')
    fresh.set_params(net.get_params())
    pi_f, v_f = fwd(fresh)
    assert np.array_equal(pi2, pi_f) and np.array_equal(v2, v_f)
    assert not np.array_equal(v1, v2)


def test_sliced_bootstrap_equals_whole_batch():
    """RolloutEngine.bootstrap(lo, hi) per environment slice (what the end-to-end loop issues as each slice's last frames
    arrive) leaves the same V(s_T), loss and update as the whole-batch bootstrap inside update()."""
    arch, A, N, T = 'NATURE', 6, 24, 2
    rng = np.random.RandomState(13)
    states = rng.randint(0, 256, (T + 1, N, 84, 84, 4)).astype(np.uint8)
    rewards = rng.choice([-1.0, 0.0, 1.0], size=(T, N)).astype(np.float32)
    outs = []
    for sliced in (False, True):
        net = G.make_net(arch, A, seed=5, math='bf16x3')
        eng = RolloutEngine(net, N, T, seed=9)
        eng.set_states(G.dev(states))
        eng.draw_uniforms()
        for t in range(T):
            eng.act(t)
        eng.rewards.copy_(G.dev(rewards))
        if sliced:
            eng.bootstrap(8, 24)          # slice 0 is left to update()
        eng.update(0.0224)
        torch.cuda.synchronize()
        outs.append((eng.boot_v.cpu().numpy(), eng.y.cpu().numpy(), eng.loss.item(), net.get_params()))
    assert np.array_equal(outs[0][0].view(np.int32), outs[1][0].view(np.int32))
    assert np.array_equal(outs[0][1].view(np.int32), outs[1][1].view(np.int32))
    assert abs(outs[0][2] - outs[1][2]) <= 1e-6 * max(1.0, abs(outs[0][2]))
    assert_close(outs[1][3], outs[0][3], 1e-6, 'post-update weights')


@pytest.mark.parametrize('b', [1, 2, 3, 37, 148 * 2 + 1, 1111])
def test_conv3_packed_tiles_equal_input_grid_tiles(monkeypatch, b):
    """conv3's forward with two whole samples per tile (Geo<G_FWD3P>: three patch copies, row groups at a 9-unit stride) against the
    input-grid enumeration (PAACB_CONV3_PACKED=0): same products, same order per output row -- but the tensor core rounds a run of
    MMAs into one accumulator differently when the run is issued in three pieces instead of one (measured: mean difference one
    fp32 ulp), so conv1 / conv2 must be the same bits and conv3 / fc / pi / v equal to 2e-5 of their scale.  Odd batches
    (zero-filled half tile), one tile, more tiles than SMs, a slice of a larger workspace."""
    gen = torch.Generator(device='cuda'); gen.manual_seed(100 + b)
    states = torch.randint(0, 256, (b, 84, 84, 4), dtype=torch.uint8, device='cuda', generator=gen)
    out = []
    cap, first = b + 5, 3
    for knob in ('0', '1'):
        monkeypatch.setenv('PAACB_CONV3_PACKED', knob)
        net = G.make_net('NATURE', 6, seed=21, math='bf16x3')
        pi = torch.full((b, 6), -1.0, device='cuda'); v = torch.full((b,), -1.0, device='cuda')
        ws = torch.zeros((net.workspace_floats(cap),), device='cuda')
        net.forward(states, pi, v, ws, ws_capacity=cap, ws_first=first)
        torch.cuda.synchronize()
        out.append((ws.view(torch.int16).clone(), pi, v))
    w0, w1 = out[0][0], out[1][0]
    off = 0
    for li, e in enumerate([20 * 20 * 32, 9 * 9 * 64, 7 * 7 * 64, 512]):
        base = off * cap * 2                    # int16 index of the layer's region: hi plane, then lo plane, e * cap each
        if li < 2:
            assert torch.equal(w0[base:base + 2 * e * cap], w1[base:base + 2 * e * cap]), 'layer %d' % (li + 1)
        else:
            def val(w):
                hi = (w[base:base + e * cap].to(torch.int32) << 16).view(torch.float32)
                lo = (w[base + e * cap:base + 2 * e * cap].to(torch.int32) << 16).view(torch.float32)
                return hi.double() + lo.double()
            v0, v1 = val(w0), val(w1)
            assert v0.abs().max() > 0 and (v0[:first * e] == 0).all() and (v1[:first * e] == 0).all()      # outside the slice: untouched
            assert (v0 - v1).abs().max() <= 2e-5 * v0.abs().max(), 'layer %d' % (li + 1)
            assert ((v0 == 0) == (v1 == 0)).float().mean() > 0.999         # the same ReLU pattern up to roundings across zero
        off += e
    assert torch.isfinite(out[1][1]).all()
    assert (out[0][1] - out[1][1]).abs().max() <= 1e-5 and (out[0][2] - out[1][2]).abs().max() <= 1e-5 * max(1.0, float(out[0][2].abs().max()))
