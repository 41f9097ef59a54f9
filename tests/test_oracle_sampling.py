"""CPU: the oracle's Philox4x32-10 against the Random123 known-answer vectors, and the uniform / inverse-CDF rules."""
import numpy as np

from oracle import sampling


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors: philox4x32 10 <counter x4> <key x2> <expected x4>
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = sampling.philox4x32_10(np.asarray([ctr], np.uint32), np.asarray([key], np.uint32))[0]
        assert [int(x) for x in got] == list(want)


def test_uniforms_are_in_unit_interval_and_stateless():
    u = sampling.uniforms(seed=3, draw=7, first_sample=0, count=4096)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    assert abs(float(u.mean()) - 0.5) < 0.02
    # slices draw what the whole batch draws; other draws / seeds differ
    assert np.array_equal(sampling.uniforms(3, 7, 1000, 96), u[1000:1096])
    assert not np.array_equal(sampling.uniforms(3, 8, 0, 4096), u)
    assert not np.array_equal(sampling.uniforms(4, 7, 0, 4096), u)


def test_inverse_cdf_rule():
    pi = np.asarray([[0.25, 0.25, 0.5], [0.0, 1.0, 0.0], [0.1, 0.2, 0.7]], np.float32)
    assert sampling.sample_actions(pi, [0.0, 0.999, 0.3]).tolist() == [0, 1, 2]
    assert sampling.sample_actions(pi, [0.25, 0.0, 0.29999]).tolist() == [1, 1, 1]
    assert sampling.sample_actions(pi, [0.4999, 0.5, 0.0999]).tolist() == [1, 1, 0]
    # a distribution check against np.random.multinomial's definition (paac.py:42-44): frequencies match pi
    rng = np.random.RandomState(0)
    p = np.asarray([[0.05, 0.15, 0.3, 0.5]], np.float32).repeat(20000, 0)
    a = sampling.sample_actions(p, rng.random_sample(20000).astype(np.float32))
    freq = np.bincount(a, minlength=4) / 20000.0
    assert np.max(np.abs(freq - p[0])) < 0.01
