"""GPU: the round-2 engine features through the C ABI --
in-kernel Philox sampling (K6) against the oracle's Philox + inverse-CDF restatement, paacb_observe_u8 (K1 + per-step
bookkeeping in one launch), the ping-pong rollout states (no copy between rollouts), CUDA-graph replay of act / update
against the eager calls, the opt-in gradient statistics, and the current-device guard."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import sampling
from oracle import preprocess as opre
from paac_b200 import _lib
from paac_b200.engine import RolloutEngine
from paac_b200.resize_tables import ROW, COL
from util import assert_close
import gpu_util as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('arch,math,A,b', [('NATURE', 'bf16x3', 6, 700), ('NIPS', 'fp32', 4, 129), ('NIPS', 'bf16x3', 18, 257)])
def test_in_kernel_philox_sampling_matches_oracle(arch, math, A, b):
    net = G.make_net(arch, A, seed=3, math=math)
    rng_state = torch.tensor([12345, 40], dtype=torch.int64, device='cuda')
    states = G.dev(np.random.RandomState(b).randint(0, 256, (b, 84, 84, 4)).astype(np.uint8))
    pi = torch.empty((b, A), device='cuda'); v = torch.empty((b,), device='cuda')
    ws = torch.empty((net.workspace_floats(b),), device='cuda')
    act = torch.full((b,), -1, dtype=torch.int32, device='cuda')
    oh = torch.full((b, A), -1.0, device='cuda')
    net.forward(states, pi, v, ws, actions=act, onehot=oh, rng=rng_state, draw=2, first_sample=1000)
    torch.cuda.synchronize()
    u = sampling.uniforms(seed=12345, draw=42, first_sample=1000, count=b)
    want = sampling.sample_actions(pi.cpu().numpy(), u)
    got = act.cpu().numpy()
    assert np.array_equal(got, want)                                     # bit-exact indices given the GPU's own pi
    assert np.array_equal(oh.cpu().numpy(), np.eye(A, dtype=np.float32)[got])      # paac.py:27
    # environment slices draw what the whole batch draws
    act2 = torch.full((b,), -1, dtype=torch.int32, device='cuda')
    cut = b // 3
    for lo, hi in ((0, cut), (cut, b)):
        net.forward(states[lo:hi], pi[lo:hi], v[lo:hi], ws, actions=act2[lo:hi], rng=rng_state, draw=2, first_sample=1000 + lo)
    torch.cuda.synchronize()
    assert np.array_equal(act2.cpu().numpy(), got)
    # paacb_rng_advance moves the draw base in stream order
    _lib.check(net._lib.paacb_rng_advance(net.ctx, _lib.ptr(rng_state), 5, G.stream()), 'paacb_rng_advance')
    net.forward(states, pi, v, ws, actions=act2, rng=rng_state, draw=2, first_sample=1000)
    torch.cuda.synchronize()
    assert rng_state.tolist() == [12345, 45]
    assert np.array_equal(act2.cpu().numpy(), sampling.sample_actions(pi.cpu().numpy(), sampling.uniforms(12345, 47, 1000, b)))
    # both sources at once is an argument error
    rc = net._lib.paacb_policy_forward_sample(net.ctx, _lib.ptr(net.params), _lib.ptr(states), b, _lib.ptr(ws), b, 0, _lib.ptr(pi),
                                              _lib.ptr(v), None, 0, 0, _lib.ptr(act), None, G.stream())
    assert rc == -1


def test_sampling_frequencies_follow_pi():
    """Distribution check (what np.random.multinomial of paac.py:42-44 is held to): 64 draws x 4096 samples of one state."""
    A, b = 6, 4096
    net = G.make_net('NATURE', A, seed=3, math='bf16x3')
    one = np.random.RandomState(0).randint(0, 256, (1, 84, 84, 4)).astype(np.uint8)
    states = G.dev(np.repeat(one, b, 0))
    pi = torch.empty((b, A), device='cuda'); v = torch.empty((b,), device='cuda')
    ws = torch.empty((net.workspace_floats(b),), device='cuda')
    act = torch.empty((b,), dtype=torch.int32, device='cuda')
    rng_state = torch.tensor([7, 0], dtype=torch.int64, device='cuda')
    counts = np.zeros(A)
    for d in range(64):
        net.forward(states, pi, v, ws, actions=act, rng=rng_state, draw=d)
        counts += np.bincount(act.cpu().numpy(), minlength=A)
    p = pi[0].cpu().numpy().astype(np.float64)
    freq = counts / counts.sum()
    assert np.max(np.abs(freq - p)) < 4.0 * np.sqrt(p.max() / counts.sum()) + 1e-3


def test_observe_u8_is_preprocess_plus_bookkeeping():
    """paacb_observe_u8: same states as paacb_preprocess_u8 with explicit reset flags (bit-exact vs the oracle built from the
    reference's FramePool / ObservationPool), and the step's rewards / episode-over flags land in their rollout row."""
    N, T, A = 37, 2, 6
    net = G.make_net('NATURE', A, seed=3, math='fp32')
    eng = RolloutEngine(net, N, T, seed=1)
    rng = np.random.RandomState(5)
    s0 = rng.randint(0, 256, (N, 84, 84, 4)).astype(np.uint8)
    frames = rng.randint(0, 256, (T, N, 4, 2, 210, 160)).astype(np.uint8)
    rewards = rng.choice([-2.0, 0.0, 1.0, 9.0], size=(T, N)).astype(np.float32)
    over = (rng.random_sample((T, N)) < 0.3).astype(np.float32)
    eng.state(0).copy_(G.dev(s0))
    want = [s0]
    for t in range(T):
        f = G.dev(frames[t])
        eng.observe_frames(t, f.data_ptr(), 4, None, G.dev(rewards[t]), G.dev(over[t]), over_is_reset=True)
        torch.cuda.synchronize()
        want.append(opre.step_states(want[-1], frames[t], over[t].astype(np.uint8), ROW, COL))
    assert np.array_equal(eng.get_states().cpu().numpy(), np.stack(want))
    assert np.array_equal(eng.rewards.cpu().numpy(), rewards) and np.array_equal(eng.over.cpu().numpy(), over)
    # without over_is_reset the flags are stored but nothing resets; slices address their own rows
    eng2 = RolloutEngine(net, N, T, seed=1)
    eng2.state(0).copy_(G.dev(s0))
    f = G.dev(frames[0])
    for lo, hi in ((0, 10), (10, N)):
        eng2.observe_frames(0, f[lo].data_ptr(), 4, None, G.dev(rewards[0]), G.dev(over[0]), lo, hi)
    torch.cuda.synchronize()
    assert np.array_equal(eng2.state(1).cpu().numpy(), opre.step_states(s0, frames[0], np.zeros(N, np.uint8), ROW, COL))
    assert np.array_equal(eng2.over[0].cpu().numpy(), over[0]) and np.array_equal(eng2.rewards[0].cpu().numpy(), rewards[0])


def test_rollouts_chain_without_a_copy():
    """s_T of a rollout IS s_0 of the next (paac.py:99-112): two consecutive rollouts see one continuous state sequence."""
    N, T, A = 8, 3, 6
    net = G.make_net('NATURE', A, seed=3, math='bf16x3')
    eng = RolloutEngine(net, N, T, seed=1)
    rng = np.random.RandomState(9)
    s = rng.randint(0, 256, (N, 84, 84, 4)).astype(np.uint8)
    eng.state(0).copy_(G.dev(s))
    zero = torch.zeros(N, device='cuda')
    seq = [s]
    for cycle in range(3):
        first_ptr = eng.state(0).data_ptr()
        for t in range(T):
            eng.act(t)
            fr = rng.randint(0, 256, (N, 1, 2, 210, 160)).astype(np.uint8)
            eng.observe_frames(t, G.dev(fr).data_ptr(), 1, None, zero, zero)
            torch.cuda.synchronize()
            seq.append(opre.step_states(seq[-1], fr, np.zeros(N, np.uint8), ROW, COL))
        assert np.array_equal(eng.get_states().cpu().numpy(), np.stack(seq[-(T + 1):]))
        last_ptr = eng.state(T).data_ptr()
        eng.update(0.01)
        assert eng.state(0).data_ptr() == last_ptr and last_ptr != first_ptr
        assert np.array_equal(eng.state(0).cpu().numpy(), seq[-1])


@pytest.mark.parametrize('arch,math,mode', [('NATURE', 'bf16x3', 'batched'), ('NATURE', 'bf16x3', 'reuse'),
                                            ('NIPS', 'bf16x3', 'stepwise'), ('NIPS', 'tf32x3', 'batched')])
def test_graph_replay_equals_eager(arch, math, mode):
    """The reference's default size (32 environments, train.py:95): act / train_forward_step / update replayed as CUDA graphs
    leave the same actions, values, returns and loss BIT FOR BIT as the eager calls (the forward path has no atomics), and the
    same parameters to fp32 summation order (the weight-gradient kernels add their partial sums with atomics; two eager runs
    differ by as much).  The learning rate changes between replays through device memory."""
    N, T, A, cycles = 32, 5, 6, 4
    rng = np.random.RandomState(2)
    s0 = rng.randint(0, 256, (N, 84, 84, 4)).astype(np.uint8)
    frames = [G.dev(rng.randint(0, 256, (N, 1, 2, 210, 160)).astype(np.uint8)) for _ in range(cycles * T)]
    rew = G.dev(rng.choice([-1.0, 0.0, 1.0], size=(cycles * T, N)).astype(np.float32))
    over = G.dev((rng.random_sample((cycles * T, N)) < 0.1).astype(np.float32))
    out = {}
    for graphs in (False, True):
        net = G.make_net(arch, A, seed=5, math=math)
        eng = RolloutEngine(net, N, T, seed=11, train_forward=mode)
        eng.state(0).copy_(G.dev(s0))
        if graphs:
            eng.enable_graphs()
        rec = []
        for c in range(cycles):
            for t in range(T):
                eng.act(t)
                if mode == 'stepwise':
                    eng.train_forward_step(t)
                eng.observe_frames(t, frames[c * T + t].data_ptr(), 1, None, rew[c * T + t], over[c * T + t])
            eng.update(0.0224 * (1.0 - 0.2 * c))
            torch.cuda.synchronize()
            rec.append((eng.actions.cpu().numpy().copy(), eng.values.cpu().numpy().copy(), eng.y.cpu().numpy().copy(),
                        float(eng.loss.item()), float(eng.norm.item()), net.get_params()))
        out[graphs] = rec
        if graphs:
            assert eng._graphs and all(k[-1] in (0, 1) for k in eng._graphs)
    for c, (e, g) in enumerate(zip(out[False], out[True])):
        if c == 0:          # same parameters going in: the forward path is bit-identical
            assert np.array_equal(e[0], g[0]) and np.array_equal(e[1].view(np.int32), g[1].view(np.int32))
            assert np.array_equal(e[2].view(np.int32), g[2].view(np.int32)) and e[3] == g[3]
        assert abs(e[3] - g[3]) <= 1e-4 * max(1.0, abs(e[3])) and abs(e[4] - g[4]) <= 1e-4 * e[4]
        # the fp32 atomics of the weight gradients land in a different order on every run: after the first update the two
        # runs differ by rounding, and a pre-activation that rounds across zero flips a ReLU (observed: a repeatable 3.6e-5
        # jump at cycle 3 in about half of the runs, eager vs eager as well) -- later cycles are held to the parity bar
        assert_close(g[5], e[5], 1e-5 if c == 0 else 1e-4, 'parameters after cycle %d, graph replay vs eager' % c, elem_max=None)
    assert not np.array_equal(out[True][0][5], out[True][-1][5])


def test_grad_stats_match_numpy(tmp_path):
    """Opt-in summaries (actor_learner.py:85-87, logger_utils.py:23-33): mean / stddev / max / min of the raw and clipped flat
    gradient and global_norm from ONE reduction pass."""
    import json
    from paac_b200.logger_utils import StatsWriter
    N, T, A = 16, 5, 6
    net = G.make_net('NATURE', A, seed=5, math='bf16x3')
    eng = RolloutEngine(net, N, T, seed=3)
    eng.set_states(G.dev(np.random.RandomState(0).randint(0, 256, (T + 1, N, 84, 84, 4)).astype(np.uint8)))
    for t in range(T):
        eng.act(t)
    eng.rewards.fill_(1.0)
    eng.forward_backward(); eng.allreduce()
    w = StatsWriter(str(tmp_path), every=1)
    st = w.gradients(160, eng, 0.01)
    w.episode(160, 21.0, 812)
    w.close()
    g = eng.grads.cpu().numpy().astype(np.float64)
    assert abs(st['mean'] - g.mean()) <= 1e-9 + 1e-6 * abs(g.mean()) and abs(st['stddev'] - g.std()) <= 1e-6 * g.std()
    assert st['max'] == g.max() and st['min'] == g.min()
    norm = np.sqrt((g * g).sum())
    assert abs(st['norm'] - norm) <= 1e-9 * norm
    eng.apply(0.01)
    torch.cuda.synchronize()
    assert abs(eng.norm.item() - norm) <= 1e-5 * norm
    recs = [json.loads(l) for l in open(w.path)]
    tags = {r['tag'] for r in recs}
    assert {'summaries/raw_gradients/mean', 'summaries/clipped_gradients/stddev', 'global_norm', 'rl/reward',
            'rl/episode_length'} <= tags
    by = {r['tag']: r['value'] for r in recs}
    k = 3.0 * min(1.0 / norm, 1.0 / 3.0)
    assert abs(by['summaries/clipped_gradients/max'] - g.max() * k) <= 1e-9 + 1e-6 * abs(g.max() * k)


def test_wrong_current_device_is_refused():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    net = G.make_net('NATURE', 6, seed=3, math='bf16x3')
    b = 4
    states = torch.zeros((b, 84, 84, 4), dtype=torch.uint8, device='cuda:0')
    pi = torch.empty((b, 6), device='cuda:0'); v = torch.empty((b,), device='cuda:0')
    ws = torch.empty((net.workspace_floats(b),), device='cuda:0')
    torch.cuda.set_device(1)
    try:
        rc = net._lib.paacb_policy_forward(net.ctx, _lib.ptr(net.params), _lib.ptr(states), b, _lib.ptr(ws), _lib.ptr(pi),
                                           _lib.ptr(v), None, None, None, None)
        assert rc == -1 and b'current CUDA device' in net._lib.paacb_last_error()
        # a second context on the other device works (kernel attributes are kept per device)
        conf = dict(name='local_learning', num_actions=6, clip_norm=3.0, clip_norm_type='global', device='/gpu:1',
                    entropy_regularisation_strength=0.02, seed=3, math='bf16x3')
        from paac_b200.policy_v_network import NaturePolicyVNetwork
        net1 = NaturePolicyVNetwork(conf)
        s1 = torch.randint(0, 256, (b, 84, 84, 4), dtype=torch.uint8, device='cuda:1')
        pi1 = torch.empty((b, 6), device='cuda:1'); v1 = torch.empty((b,), device='cuda:1')
        ws1 = torch.empty((net1.workspace_floats(b),), device='cuda:1')
        net1.forward(s1, pi1, v1, ws1)
        torch.cuda.synchronize(1)
        torch.cuda.set_device(0)
        net.forward(s1.to('cuda:0'), pi, v, ws)
        torch.cuda.synchronize(0)
        assert torch.equal(pi.cpu(), pi1.cpu()) and torch.equal(v.cpu(), v1.cpu())
    finally:
        torch.cuda.set_device(0)
