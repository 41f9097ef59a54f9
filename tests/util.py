"""Shared test helpers.  Tolerance convention for floating-point parity (north_star: "within 1e-4 relative"):
a tensor x matches its reference r when  max|x - r| <= rel * max|r| + abs_floor  (relative to the tensor's
scale; element-wise relative error is meaningless for entries that cancel to ~0)."""
import numpy as np

REL = 1e-4


def rel_err(x, r):
    x = np.asarray(x, np.float64)
    r = np.asarray(r, np.float64)
    assert x.shape == r.shape, (x.shape, r.shape)
    scale = np.max(np.abs(r)) if r.size else 0.0
    return (np.max(np.abs(x - r)) if r.size else 0.0) / max(scale, 1e-30)


def assert_close(x, r, rel=REL, what=''):
    e = rel_err(x, r)
    assert e <= rel, '%s: scaled max error %.3e > %.1e' % (what, e, rel)
    return e
