"""CPU: the oracle's observation pipeline against golden vectors produced by the reference's own
FramePool / ObservationPool (+ Pillow NEAREST) -- tests/golden/preprocess_golden.npz, oracle/make_golden.py."""
import os

import numpy as np
import pytest

from oracle import preprocess as opre
from oracle.make_golden import gen_frames


@pytest.fixture(scope='module')
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, 'preprocess_golden.npz'))


def test_tables_match_installed_pillow(gold):
    row, col = opre.pillow_nearest_tables()
    assert (row == gold['row_tab']).all() and (col == gold['col_tab']).all()
    # ROW = floor((y+.5)*2.5); COL = floor((x+.5)*160/84) except two tie columns
    assert (row == np.floor((np.arange(84) + 0.5) * 2.5)).all()
    f = np.floor((np.arange(84) + 0.5) * 160 / 84).astype(int)
    diff = np.nonzero(col != f)[0]
    assert list(diff) == [52, 73] and col[52] == 99 and col[73] == 139


def test_product_tables_equal_oracle_tables(gold):
    from paac_b200.resize_tables import ROW, COL
    assert (ROW == gold['row_tab']).all() and (COL == gold['col_tab']).all()


def test_step_states_reproduces_reference_sequence(gold):
    n_envs, steps = int(gold['n_envs']), int(gold['steps'])
    frames = gen_frames(int(gold['seed']), n_envs, int(gold['n_pairs']))
    used, resets, states = gold['used'], gold['resets'], gold['states']
    row, col = gold['row_tab'], gold['col_tab']
    prev = np.zeros((n_envs, 84, 84, 4), np.uint8)
    for t in range(steps + 1):
        slots = np.zeros((n_envs, 4, 2, 210, 160), np.uint8)
        for e in range(n_envs):
            k = 4 if resets[t, e] else 1
            for j in range(k):
                slots[e, j] = frames[e, used[t, e, j]]
        nxt = opre.step_states(prev, slots, resets[t], row, col)
        assert (nxt == states[t]).all(), 'step %d' % t
        prev = nxt
    assert resets[1:].sum() == 4      # the fixture exercises resets, including back-to-back ones


def test_observation_ring_matches_shift_semantics():
    ring = opre.ObservationRing()
    rng = np.random.RandomState(0)
    planes = [rng.randint(0, 256, (84, 84)).astype(np.uint8) for _ in range(9)]
    for i, p in enumerate(planes):
        ring.new_observation(p)
        obs = ring.get_pooled_observations()
        assert (obs[:, :, 3] == p).all()
        if i >= 3:
            for k in range(4):
                assert (obs[:, :, k] == planes[i - 3 + k]).all()


def test_max_pool_is_elementwise_and_commutative():
    rng = np.random.RandomState(1)
    pair = rng.randint(0, 256, (2, 210, 160)).astype(np.uint8)
    row, col = opre.pillow_nearest_tables()
    a = opre.process_frame_pool(pair, row, col)
    b = opre.process_frame_pool(pair[::-1], row, col)
    assert (a == b).all()
    assert (a == np.maximum(pair[0], pair[1])[row[:, None], col[None, :]]).all()


def test_row_table_is_two_uniform_pitches():
    """K1's copy pipeline lands the selected frame rows with two rank-3 TMA boxes (csrc/preprocess.cu): that needs Pillow's
    NEAREST row table for 210 -> 84 to be row(2k) = 5k + 1, row(2k + 1) = 5k + 3 -- each parity a uniform 5-row (800-byte)
    pitch that also runs across frames (210 = 42 * 5) -- and the last selected row of a frame to lie inside it.  (The launcher
    checks the pattern on the tables it is given and falls back to the CTA-per-environment kernel otherwise.)"""
    from paac_b200.resize_tables import ROW, COL
    rows = [int(r) for r in ROW]
    assert len(rows) == 84 and len(COL) == 84
    assert rows[0::2] == [5 * k + 1 for k in range(42)]
    assert rows[1::2] == [5 * k + 3 for k in range(42)]
    assert 210 == 42 * 5 and max(rows) < 210
    # the same table from its definition: floor((i + 0.5) * 210 / 84), in integers
    assert rows == [((2 * i + 1) * 210) // 168 for i in range(84)]
