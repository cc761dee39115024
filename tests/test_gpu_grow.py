"""Growing training set (include/gpb200.h gpb_gpr_grow_*; GP_parameter_fit.py:61-63) against full refits of the oracle.

The reference refits from scratch on every prefix; the device path extends the stored factor.  Both must agree:
NLML to rel 1e-8 and predictions to rel 1e-9 of the prediction scale, at sizes that start, end and stay inside
128-tiles and across several of them."""
import numpy as np
import pytest

from oracle import gpr_oracle

pytestmark = pytest.mark.gpu


def data(n, d, seed):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    w = rng.standard_normal(d)
    y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
    Z = rng.random((300, d))
    log_hyp = np.log([0.5] * d + [1.0, 0.1])
    return X, y, Z, log_hyp


def natural(log_hyp):
    from gptest_b200.sweep import natural_params
    return natural_params(log_hyp)[0]


@pytest.mark.parametrize('cuts', [
    [5, 10, 15, 20, 25],                       # the script's five-point steps inside the first tile
    [100, 128, 129, 255, 256, 300],            # ending on, just after and before tile boundaries
    [1, 2, 400, 401, 1000],                    # single points and jumps over several tiles
    [640, 641, 1300],
    [300, 301, 306, 338, 341, 383, 384, 385, 390, 512, 517, 518],   # thin appends (substitution path) and tile crossings
])
def test_append_matches_full_refit(handle, cuts):
    X, y, Z, lh = data(cuts[-1], 3, seed=len(cuts) + cuts[-1])
    handle.grow_begin(natural(lh), 3, capacity=cuts[-1])
    n0 = 0
    for n1 in cuts:
        v = handle.grow_append(X[n0:n1], y[n0:n1])
        assert handle.grow_size() == n1
        ref = gpr_oracle.nlml_chol(lh, X[:n1], y[:n1])
        assert abs(v - ref) <= 1e-8 * max(1.0, abs(ref)), (n1, v, ref)
        fz, cov = handle.grow_predict(Z)
        rf, rc = gpr_oracle.predict_chol(lh, X[:n1], y[:n1], Z)
        assert np.abs(fz - rf).max() <= 1e-9 * max(1.0, np.abs(rf).max()), n1
        assert np.abs(cov - rc).max() <= 1e-9 * max(1.0, np.abs(rc).max()), n1
        n0 = n1


def test_append_equals_single_call_and_mean(handle):
    X, y, Z, lh = data(700, 2, seed=3)
    kh = natural(lh)
    handle.set_train(X, y)
    full = handle.gpr_nlml(kh, mean=0.3)
    pf, pc = handle.gpr_predict(kh, Z, mean=0.3)
    handle.grow_begin(kh, 2, capacity=1000, mean=0.3)
    for a, b in ((0, 333), (333, 334), (334, 700)):
        v = handle.grow_append(X[a:b], y[a:b])
    assert abs(v - full) <= 1e-11 * abs(full)
    # other calls on the handle do not disturb the stored factor
    handle.gpr_nlml(kh * 1.1)
    fz, cov = handle.grow_predict(Z)
    assert np.abs(fz - pf).max() <= 1e-10 * max(1.0, np.abs(pf).max())
    assert np.abs(cov - pc).max() <= 1e-10


def test_predict_more_points_than_one_pass(handle):
    X, y, _, lh = data(500, 2, seed=5)
    rng = np.random.default_rng(1)
    Z = rng.random((5000, 2))                     # > 2 sweeps of 2048 test rows, ragged tail
    handle.grow_begin(natural(lh), 2, capacity=500)
    handle.grow_append(X, y)
    fz, cov = handle.grow_predict(Z)
    rf, rc = gpr_oracle.predict_chol(lh, X, y, Z)
    assert np.abs(fz - rf).max() <= 1e-9 * max(1.0, np.abs(rf).max())
    assert np.abs(cov - rc).max() <= 1e-9


def test_errors(handle):
    X, y, _, lh = data(40, 2, seed=7)
    handle.grow_begin(natural(lh), 2, capacity=32)
    handle.grow_append(X[:30], y[:30])
    with pytest.raises(RuntimeError, match='capacity'):
        handle.grow_append(X[30:40], y[30:40])
    # duplicated points with no noise: not positive definite -> LinAlgError like np.linalg.cholesky (GPr.py:62)
    kh = natural(np.log([0.5, 0.5, 1.0, 1e-30]))
    handle.grow_begin(kh, 2, capacity=64)
    handle.grow_append(X[:10], y[:10])
    with pytest.raises(np.linalg.LinAlgError):
        handle.grow_append(np.vstack([X[:5], X[:5]]), np.r_[y[:5], y[:5]])
    with pytest.raises(RuntimeError, match='invalid'):
        handle.grow_predict(X[:3])
