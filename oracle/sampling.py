"""TEST INFRASTRUCTURE ONLY (never imported by paac_b200/).  CPU restatement of K6, the categorical sampling of
paac.py:34-45, as the product defines it:

  * reference: ``np.random.multinomial(1, p - epsneg)`` per environment (paac.py:42-44) -- NumPy's global Mersenne
    Twister, which cannot be reproduced call by call on a GPU.  Parity is therefore defined on the DISTRIBUTION and on
    the decision rule: inverse CDF on a uniform u in [0, 1) with an fp32 running sum in action order, last bucket
    open-ended; given (pi, u) the action index is checked bit-exact.
  * the uniforms: either injected by the caller, or Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers:
    as easy as 1, 2, 3", SC'11 -- the counter-based generator behind curand / torch.cuda), counter = (sample index lo,
    sample index hi, draw lo, draw hi), key = (seed lo, seed hi), u = (first output word >> 8) * 2^-24.
    Pinned by the Random123 known-answer vectors (tests/test_oracle_sampling.py).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: uint32 [..., 4], key: uint32 [..., 2] -> uint32 [..., 4] (vectorised over leading dims)."""
    c = [np.asarray(ctr[..., i], np.uint32).copy() for i in range(4)]
    k0 = np.asarray(key[..., 0], np.uint32).copy()
    k1 = np.asarray(key[..., 1], np.uint32).copy()
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = M0 * c[0].astype(np.uint64)
            p1 = M1 * c[2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return np.stack(c, axis=-1)


def uniforms(seed, draw, first_sample, count):
    """The uniforms paacb_policy_forward_sample uses for samples first_sample .. first_sample + count of draw index `draw`."""
    i = np.arange(first_sample, first_sample + count, dtype=np.uint64)
    ctr = np.empty((count, 4), np.uint32)
    ctr[:, 0] = (i & MASK).astype(np.uint32)
    ctr[:, 1] = (i >> np.uint64(32)).astype(np.uint32)
    d = np.uint64(draw)
    ctr[:, 2] = np.uint32(d & MASK)
    ctr[:, 3] = np.uint32(d >> np.uint64(32))
    s = np.uint64(seed)
    key = np.empty((count, 2), np.uint32)
    key[:, 0] = np.uint32(s & MASK)
    key[:, 1] = np.uint32(s >> np.uint64(32))
    x = philox4x32_10(ctr, key)[:, 0]
    return (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def sample_actions(pi, u):
    """Inverse CDF with an fp32 running sum in index order; the last action absorbs the remainder (App. E.4)."""
    pi = np.asarray(pi, np.float32)
    u = np.asarray(u, np.float32)
    b, A = pi.shape
    act = np.full(b, A - 1, np.int32)
    c = np.zeros(b, np.float32)
    done = np.zeros(b, bool)
    for j in range(A - 1):
        c = (c + pi[:, j]).astype(np.float32)
        hit = (~done) & (u < c)
        act[hit] = j
        done |= hit
    return act
