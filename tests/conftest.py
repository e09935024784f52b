import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def dense(rows, cols, vals, shape):
    m = np.zeros(tuple(int(s) for s in shape), np.float32)
    m[rows, cols] = vals
    return m


def assert_parity(a, b, rel=1e-4, what=""):
    """SURVEY.md §8(d) tolerance: max|a-b| <= rel*max|b| AND allclose(rtol=rel, atol=rel/10*max|b|)."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, "%s shape %s vs %s" % (what, a.shape, b.shape)
    assert np.isfinite(a).all() == np.isfinite(b).all(), what + " non-finite mismatch"
    peak = float(np.abs(b).max()) if b.size else 0.0
    err = float(np.abs(a - b).max()) if b.size else 0.0
    assert err <= rel * peak + 1e-30, "%s max abs err %.3e > %.0e * peak %.3e" % (what, err, rel, peak)
    assert np.allclose(a, b, rtol=rel, atol=rel / 10 * peak), "%s allclose(rtol=%g, atol=%g) failed; worst %.3e" % (
        what, rel, rel / 10 * peak, float((np.abs(a - b) - rel * np.abs(b)).max()))


def branch_cut(X, tol=1e-4):
    """Elements whose phase is decided by rounding noise: negative real part with a relatively
    tiny imaginary part (angle = +pi or -pi by the sign of ~1e-7 noise), or |X| ~ 0.
    Frame 0 of a centre-padded STFT is a symmetric frame, so (almost) every bin of it sits here."""
    X = np.asarray(X)
    a = np.abs(X)
    return ((np.abs(X.imag) <= tol * a) & (X.real <= 0)) | (a <= tol * a.max())


def if_mask(X, method="forward", tol=1e-4):
    """True where an IF output row is well conditioned: neither of the two frames it differences
    sits on the branch cut (row 0 / last row carry the raw phase of one frame)."""
    bad = branch_cut(X, tol)
    m = ~bad
    if method == "forward":
        m[..., 1:, :] &= ~bad[..., :-1, :]
    elif method == "backward":
        m[..., :-1, :] &= ~bad[..., 1:, :]
    else:
        m[..., 1:-1, :] = ~bad[..., 2:, :] & ~bad[..., :-2, :]
    return m


def unwrap_mask(X, tol=1e-4):
    """True where an unwrapped phase is well conditioned: no frame at or before it (same bin)
    sits on the branch cut — a +-pi flip there shifts every later frame by 2 pi."""
    return ~(np.cumsum(branch_cut(X, tol), axis=-2) > 0)


@pytest.fixture
def golden():
    return load_golden
