"""Shared test helpers.  Tolerance convention for floating-point parity (north_star: "within 1e-4 relative"):

  * scaled error:       max|x - r| <= rel * max|r|                      (relative to the tensor's scale)
  * element-wise error: over the SIGNIFICANT entries, |r| > 1e-3 * max|r| (entries that cancel to ~0 have no meaningful
                        relative error), the median of |x - r| / |r| is held to `rel` as well and the maximum to
                        `elem_max` (default 5e-2: an entry a thousand times smaller than the tensor's largest may carry
                        an absolute error of 5e-5 of the scale).  Both figures are printed (pytest -s / on failure)
                        and collected in REPORT so a run can dump them (tools/parity_report.py).
"""
import numpy as np

REL = 1e-4
ELEM_MAX = 5e-2
SIGNIFICANT = 1e-3
REPORT = []


def rel_err(x, r):
    x = np.asarray(x, np.float64)
    r = np.asarray(r, np.float64)
    assert x.shape == r.shape, (x.shape, r.shape)
    scale = np.max(np.abs(r)) if r.size else 0.0
    return (np.max(np.abs(x - r)) if r.size else 0.0) / max(scale, 1e-30)


def elem_err(x, r, significant=SIGNIFICANT):
    """(max, median, count) of the element-wise relative error over entries with |r| > significant * max|r|."""
    x = np.asarray(x, np.float64).reshape(-1)
    r = np.asarray(r, np.float64).reshape(-1)
    if not r.size or np.max(np.abs(r)) < 1e-30:        # a tensor that is zero to fp32 has no relative error to speak of
        return 0.0, 0.0, 0
    m = np.abs(r) > significant * np.max(np.abs(r))
    if not m.any():
        return 0.0, 0.0, 0
    e = np.abs(x[m] - r[m]) / np.abs(r[m])
    return float(e.max()), float(np.median(e)), int(m.sum())


def assert_close(x, r, rel=REL, what='', elem_max=ELEM_MAX, elem_median=None):
    e = rel_err(x, r)
    emax, emed, n = elem_err(x, r)
    REPORT.append(dict(what=what, scaled=e, elem_max=emax, elem_median=emed, significant=n, rel=rel))
    print('parity %-40s scaled %.2e (<= %.1e)  element-wise over %d entries: median %.2e  max %.2e' % (what, e, rel, n, emed, emax))
    assert e <= rel, '%s: scaled max error %.3e > %.1e' % (what, e, rel)
    if elem_max is not None:
        assert emax <= max(elem_max, 1e3 * rel * 0.5), '%s: element-wise relative error %.3e > %.1e over %d significant entries' % (
            what, emax, elem_max, n)
    med_bound = rel if elem_median is None else elem_median
    assert emed <= med_bound, '%s: median element-wise relative error %.3e > %.1e' % (what, emed, med_bound)
    return e
