"""GPU: the update path at BASELINE.json's full size (cfg3: NatureNetwork, 4096 envs, t_max 5, B = 20,480).
  * against the oracle: 64 random samples of the benchmarked bf16x3 cycle -- every layer's activations, pi and v -- vs the
    fp64 restatement (oracle.network.forward), and the FULL-size gradient / pi / v vs the fp32 SIMT anchor (itself
    oracle-checked at small sizes) run on the same 20,480-sample batch;
  * size-independent properties: schedule invariance of the forward (bit-exact), determinism, linearity of the backward,
    closed forms of the returns recurrence and of the first RMSProp step, evaluated with torch ops on the GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import network
from paac_b200 import _lib
from paac_b200.engine import RolloutEngine
from util import assert_close
import gpu_util as G

pytestmark = pytest.mark.gpu

ARCH, A, N, T = 'NATURE', 6, 4096, 5


@pytest.fixture(scope='module')
def rollout():
    gen = torch.Generator(device='cuda'); gen.manual_seed(7)
    states = torch.randint(0, 256, (T + 1, N, 84, 84, 4), dtype=torch.uint8, device='cuda', generator=gen)
    u = torch.rand((T, N), device='cuda', generator=gen)
    rewards = torch.where(u < 0.05, -3.0, torch.where(u > 0.95, 2.0, 0.0)).float()          # clipped to [-1, 1] by K7
    over = (torch.rand((T, N), device='cuda', generator=gen) < 0.02).float()
    return states, rewards, over


def run_cycle(mode, rollout, lr=0.0224, scale=1.0):
    states, rewards, over = rollout
    net = G.make_net(ARCH, A, seed=5, math='bf16x3')
    eng = RolloutEngine(net, N, T, seed=9, train_forward=mode)
    eng.set_states(states)
    eng.draw_uniforms()
    for t in range(T):
        eng.act(t)
        if mode == 'stepwise':
            eng.train_forward_step(t, 0, 1000)          # ragged slices of the step
            eng.train_forward_step(t, 1000, N)
    eng.rewards.copy_(rewards * scale); eng.over.copy_(over)
    p0 = net.params.clone()
    eng.update(lr, roll=False)           # keep this rollout's states addressable for the checks below
    torch.cuda.synchronize()
    return net, eng, p0


def test_forward_schedules_bit_identical_at_full_size(rollout):
    ref = None
    for mode in ('batched', 'stepwise', 'reuse'):
        net, eng, _ = run_cycle(mode, rollout)
        got = (eng.fwd_ws.view(torch.int32).clone(), eng.pi.clone(), eng.v.clone(), eng.actions.clone(), eng.y.clone())
        if ref is None:
            ref = got
        else:
            for a, b in zip(ref, got):
                assert torch.equal(a, b), mode
        del net, eng
        torch.cuda.empty_cache()


def test_returns_closed_form_and_determinism_at_full_size(rollout):
    states, rewards, over = rollout
    net, eng, _ = run_cycle('batched', rollout)
    # y_t = clip(r_t) + gamma * (1 - over_t) * y_{t+1}, y_T = V(s_T): float64 recurrence rounded once (paac.py:144-149)
    R = eng.boot_v.double()
    want = torch.empty((T, N), dtype=torch.float64, device='cuda')
    for t in range(T - 1, -1, -1):
        R = rewards[t].clamp(-1, 1).double() + 0.99 * R * (1.0 - over[t].double())
        want[t] = R
    assert torch.equal(eng.y.view(T, N), want.float())
    assert torch.equal(eng.adv.view(T, N), (want - eng.values.double()).float())
    # a second cycle from the same inputs: same actions / values bit for bit (the forward has no atomics)
    net2, eng2, _ = run_cycle('batched', rollout)
    assert torch.equal(eng.actions, eng2.actions) and torch.equal(eng.values, eng2.values) and torch.equal(eng.boot_v, eng2.boot_v)


def test_backward_is_linear_in_the_head_gradients_at_full_size(rollout):
    """paacb_backward(c * dlogits, c * dv) == c * paacb_backward(dlogits, dv): c = 4 is exact in every fp32 / bf16 operation
    of the path, so only the order of the fp32 atomics differs."""
    net, eng, _ = run_cycle('batched', rollout)
    B = N * T
    p = _lib.ptr
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flat = eng.flat_states
    # two backward passes over the cycle's own workspace under the same (post-update) parameters: dlogits, dv and 4 x them
    dl4, dv4 = eng.dlogits * 4.0, eng.dv * 4.0
    _lib.check(net._lib.paacb_backward(net.ctx, p(net.params), p(flat), B, p(eng.fwd_ws), p(dl4), p(dv4), p(eng.bwd_ws),
                                       p(eng.grads), st), 'paacb_backward')
    torch.cuda.synchronize()
    g4 = eng.grads.clone()
    _lib.check(net._lib.paacb_backward(net.ctx, p(net.params), p(flat), B, p(eng.fwd_ws), p(eng.dlogits), p(eng.dv),
                                       p(eng.bwd_ws), p(eng.grads), st), 'paacb_backward')
    torch.cuda.synchronize()
    g1b = eng.grads.clone()
    scale = g1b.abs().max()
    assert scale > 0 and torch.isfinite(g4).all()
    assert (g4 - 4.0 * g1b).abs().max() <= 2e-5 * 4.0 * scale


def test_first_rmsprop_step_closed_form_at_full_size(rollout):
    """From the TF slot initialisation (ms = 1, mom = 0): var -= lr * g_c / sqrt(1 + 0.01 (g_c^2 - 1) + 0.1) with
    g_c = g * clip / max(norm, clip)  (SURVEY 8c / App. B)."""
    net, eng, p0 = run_cycle('batched', rollout)
    g = eng.grads.double()
    norm = torch.sqrt((g * g).sum())
    assert abs(eng.norm.item() - norm.item()) <= 1e-5 * norm.item()
    gc = (eng.grads * (3.0 * min(1.0 / eng.norm.item(), 1.0 / 3.0))).float()
    ms = 1.0 + (gc * gc - 1.0) * (1.0 - 0.99)
    want = p0 - (gc * 0.0224) / torch.sqrt(ms + 0.1)
    err = (net.params - want).abs().max().item()
    assert err <= 1e-6 * want.abs().max().item() + 1e-9
    assert torch.allclose(eng.ms, ms, rtol=1e-6, atol=1e-9)


def test_full_size_cycle_vs_oracle_and_fp32_anchor(rollout):
    """The benchmarked configuration itself (cfg3, bf16x3, B = 20,480), not a scaled-down stand-in."""
    net, eng, p0 = run_cycle('batched', rollout)
    B = N * T
    flat = eng.flat_states
    p0_np = p0.cpu().numpy()
    params = network.unflatten_params(p0_np, ARCH, A)
    # (1) 64 random samples of the training batch against the fp64 oracle
    idx = np.sort(np.random.RandomState(1).choice(B, 64, replace=False))
    tidx = torch.from_numpy(idx).cuda()
    ref = network.forward(params, flat[tidx].cpu().numpy(), ARCH, dtype=torch.float64, keep=True)
    layers = [t[tidx].cpu().numpy() for t in net.layer_tensors(eng.fwd_ws, B)]
    for i, a in enumerate(ref['acts']):
        assert_close(layers[i], a.numpy(), 1e-4, 'full-size conv%d activation (64 samples)' % (i + 1))
    assert_close(layers[-1], ref['h'].numpy(), 1e-4, 'full-size hidden fc (64 samples)')
    assert_close(eng.pi[tidx].cpu().numpy(), ref['pi'].numpy(), 1e-4, 'full-size pi (64 samples)')
    assert_close(eng.v[tidx].cpu().numpy(), ref['v'].numpy(), 1e-4, 'full-size v (64 samples)')
    del layers
    # (2) the whole batch: forward outputs against the fp32 SIMT anchor, and the flat gradient of BOTH arithmetic modes
    # against the fp64 arbiter evaluated with torch on the GPU (gpu_util.fp64_masked_grads_cuda: same dlogits / dv, each
    # implementation's own ReLU masks -- the convention of every backward parity test)
    net32 = G.make_net(ARCH, A, seed=5, math='fp32')
    net32.set_params(p0_np)
    pi32 = torch.empty((B, A), device='cuda'); v32 = torch.empty((B,), device='cuda')
    ws32 = torch.empty((net32.workspace_floats(B),), device='cuda')
    net32.forward(flat, pi32, v32, ws32)
    bws32 = torch.empty((int(net32._lib.paacb_backward_workspace_floats(net32.ctx, B)),), device='cuda')
    g32 = torch.empty((net32.param_count,), device='cuda')
    p = _lib.ptr
    _lib.check(net32._lib.paacb_backward(net32.ctx, p(net32.params), p(flat), B, p(ws32), p(eng.dlogits), p(eng.dv), p(bws32),
                                         p(g32), C.c_void_p(torch.cuda.current_stream().cuda_stream)), 'paacb_backward')
    torch.cuda.synchronize()
    assert_close(eng.pi.cpu().numpy(), pi32.cpu().numpy(), 1e-4, 'full-size pi vs fp32 anchor')
    assert_close(eng.v.cpu().numpy(), v32.cpu().numpy(), 1e-4, 'full-size v vs fp32 anchor')
    failures = []
    for label, nn, ws, g in (('bf16x3', net, eng.fwd_ws, eng.grads), ('fp32 anchor', net32, ws32, g32)):
        acts = nn.layer_tensors(ws, B)
        want = G.fp64_masked_grads_cuda(ARCH, A, p0_np, flat, acts, eng.dlogits, eng.dv)
        del acts
        got = network.unflatten_params(g.cpu().numpy(), ARCH, A)
        for name, _, _ in network.param_specs(ARCH, A):
            try:
                assert_close(got[name], want[name], 1e-4, 'full-size gradient %s [%s vs fp64]' % (name, label))
            except AssertionError as e:
                failures.append(str(e))
    assert not failures, '\n'.join(failures)
