"""Shared helpers for the parity tests."""
import numpy as np


def rel_err(a, ref):
    """max|a-ref| / max|ref|  -- the 1e-4 contract of BASELINE.json's north_star (hard part 2)."""
    a = np.asarray(a, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    denom = np.abs(ref).max()
    return float(np.abs(a - ref).max() / (denom if denom > 0 else 1.0))


def assert_close(a, ref, tol=1e-4, what=""):
    a = np.asarray(a)
    ref = np.asarray(ref)
    assert a.shape == ref.shape, "%s: shape %s vs %s" % (what, a.shape, ref.shape)
    assert np.isfinite(a).all(), "%s: non-finite values" % what
    e = rel_err(a, ref)
    assert e <= tol, "%s: max|d|/max|ref| = %.3e > %.1e" % (what, e, tol)
    rms = float(np.sqrt(np.mean(np.asarray(ref, dtype=np.float64) ** 2)))
    assert np.allclose(a, ref, rtol=tol, atol=tol * max(rms, 1e-30)), "%s: allclose(rtol, atol=tol*rms) failed" % what
