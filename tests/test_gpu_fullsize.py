"""GPU: the update path at BASELINE.json's full size (cfg3: NatureNetwork, 4096 envs, t_max 5, B = 20,480) through
size-independent properties -- the CPU oracle takes minutes at this size, so the checks are: schedule invariance of the
forward (bit-exact), determinism, linearity of the backward, closed forms of the returns recurrence and of the first
RMSProp step, all evaluated with torch ops on the GPU."""
import ctypes as C

import numpy as np
import pytest
import torch

from paac_b200 import _lib
from paac_b200.engine import RolloutEngine
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
    eng.states.copy_(states)
    eng.draw_uniforms()
    for t in range(T):
        eng.act(t)
        if mode == 'stepwise':
            eng.train_forward_step(t, 0, 1000)          # ragged slices of the step
            eng.train_forward_step(t, 1000, N)
    eng.rewards.copy_(rewards * scale); eng.over.copy_(over)
    p0 = net.params.clone()
    eng.update(lr)
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
    flat = eng.states[:T].view((B, 84, 84, 4))
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
