"""CPU: known-answer tests for the oracle's returns / clip / RMSProp / lr / sampling / closed-form gradients
(SURVEY 8c "known-answer material to generate")."""
import numpy as np
import torch

from oracle import network, update
from util import assert_close


def test_nstep_returns_hand_cases():
    T, N, g = 5, 6, 0.99
    r = np.zeros((T, N)); over = np.zeros((T, N)); v = np.zeros((T, N), np.float32)
    boot = np.full(N, 2.0, np.float32)
    for n in range(5):
        over[n, n] = 1.0              # env n terminates at step n; env 5 never
    r[:, :] = 1.0
    r[2, 3] = 7.0                     # clipped to 1
    r[1, 4] = -3.0                    # clipped to -1
    y, adv = update.nstep_returns(r, over, v, boot, g)
    for n in range(6):
        R = 2.0
        for t in reversed(range(T)):
            rc = max(-1.0, min(1.0, r[t, n]))
            R = rc + g * R * (1.0 - over[t, n])
            assert y[t, n] == np.float32(R)
    assert y[4, 4] == 1.0 and y[0, 0] == 1.0         # terminal step: no bootstrap leaks through
    assert (adv == y).all()


def test_clip_identity_and_rmsprop_first_step():
    rng = np.random.RandomState(0)
    gs = [rng.randn(7, 5).astype(np.float32) * 3, rng.randn(11).astype(np.float32)]
    clipped, norm = update.clip_by_global_norm(gs, 3.0)
    flat = np.concatenate([g.reshape(-1) for g in gs]).astype(np.float64)
    assert abs(float(norm) - np.sqrt((flat ** 2).sum())) < 1e-5 * float(norm)
    scale = 3.0 / max(float(norm), 3.0)               # App. C: clip / max(norm, clip)
    assert_close(np.concatenate([c.reshape(-1) for c in clipped]), flat * scale, 1e-6)
    small = [g * np.float32(1e-3) for g in gs]
    c2, n2 = update.clip_by_global_norm(small, 3.0)
    assert_close(c2[0], small[0], 1e-7)                # below the threshold: untouched
    g = clipped[0]
    var = np.zeros_like(g); ms = np.ones_like(g); mom = np.zeros_like(g)
    nv, nms, nmom = update.rmsprop_apply(var, ms, mom, g, 0.0224, 0.99, 0.1)
    expect = 0.0224 * g.astype(np.float64) / np.sqrt(1 + 0.01 * (g.astype(np.float64) ** 2 - 1) + 0.1)
    assert_close(-nv, expect, 1e-5)


def test_get_lr():
    assert update.get_lr(0, 0.0224, 80000000) == 0.0224
    assert abs(update.get_lr(40000000, 0.0224, 80000000) - 0.0112) < 1e-12
    assert update.get_lr(80000000, 0.0224, 80000000) == 0.0
    assert update.get_lr(80000001, 0.0224, 80000000) == 0.0


def test_sampling_inverse_cdf_edges_and_distribution():
    pi = np.asarray([[0.2, 0.3, 0.5]] * 6, np.float32)
    u = np.asarray([0.0, 0.19999, 0.2, 0.4999, 0.5, 0.999999], np.float32)
    assert list(update.sample_actions(pi, u)) == [0, 0, 1, 1, 2, 2]
    rng = np.random.RandomState(0)
    p = np.asarray([0.05, 0.1, 0.15, 0.2, 0.25, 0.25], np.float32)
    n = 200000
    a = update.sample_actions(np.tile(p, (n, 1)), rng.random_sample(n).astype(np.float32))
    freq = np.bincount(a, minlength=6) / n
    ref = np.random.RandomState(1).multinomial(n, p.astype(np.float64) - np.finfo(np.float32).epsneg) / n   # paac.py:42-44
    assert np.abs(freq - p).max() < 5e-3 and np.abs(freq - ref).max() < 7e-3


def test_closed_form_matches_autograd_fp64():
    rng = np.random.RandomState(3)
    b, A = 17, 6
    logits = rng.randn(b, A) * 2
    v = rng.randn(b); adv = rng.randn(b); tgt = rng.randn(b)
    acts = rng.randint(0, A, b)
    zt = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    vt = torch.tensor(v, dtype=torch.float64, requires_grad=True)
    onehot = torch.nn.functional.one_hot(torch.as_tensor(acts), A).to(torch.float64)
    loss = network.a2c_loss(torch.softmax(zt, 1), vt, onehot, torch.tensor(adv), torch.tensor(tgt), 0.02)
    loss.backward()
    l2, dz, dv = network.closed_form_head_grads(logits, v, acts, adv, tgt, 0.02)
    assert abs(l2 - float(loss)) < 1e-12
    assert np.abs(dz - zt.grad.numpy()).max() < 1e-14 and np.abs(dv - vt.grad.numpy()).max() < 1e-14


def test_nature_forward_shapes_and_flatten_order():
    A = 6
    p = network.init_params('NATURE', A, 3)
    st = np.random.RandomState(0).randint(0, 256, (3, 84, 84, 4)).astype(np.uint8)
    out = network.forward(p, st, 'NATURE', keep=True)
    assert [tuple(a.shape) for a in out['acts']] == [(3, 20, 20, 32), (3, 9, 9, 64), (3, 7, 7, 64)]
    assert out['h'].shape == (3, 512) and out['pi'].shape == (3, A)
    assert np.allclose(out['pi'].sum(1).numpy(), 1, atol=1e-6)
