"""The GPy-shaped layer used by GP_parameter_fit.py (gptest_b200/gpy_compat.py) against the CPU oracle.

GPy itself is absent (parity unpinned); the arithmetic is checked: likelihood, gradient-driven optimisation,
restarts, prefix refits and dense-grid prediction on the script's own data recipe (GP_parameter_fit.py:9-33,47-63)."""
import math

import numpy as np
import pytest

from oracle import gpr_oracle

pytestmark = pytest.mark.gpu
MEAN_VALUE = 3


def explore_cost_function(a, b):
    """GP_parameter_fit.py:10-20 (the data recipe of the script)."""
    cost = MEAN_VALUE
    cost += 10 * math.exp(-math.sqrt((a - 40) ** 2 + (b - 40) ** 2) / 16)
    cost += 7 * math.exp(-math.sqrt((a - 10) ** 2 + (b - 90) ** 2) / 12)
    cost += 4 * math.exp(-math.sqrt((a - 80) ** 2 + (b - 60) ** 2) / 32)
    cost += 7 * math.exp(-math.sqrt((a + 20) ** 2 + (b - 50) ** 2) / 32)
    cost += 7 * math.exp(-math.sqrt((a - 120) ** 2 + (b - 50) ** 2) / 32)
    cost += 12 * math.exp(-math.sqrt((a - 80) ** 2 + (b - 20) ** 2) / 8)
    cost += 5 * math.exp(-math.sqrt((a - 60) ** 2 + (b - 80) ** 2) / 10)
    cost += 3 * math.exp(-math.sqrt((a - 90) ** 2 + (b - 90) ** 2) / 20)
    return cost


def script_data(n=200, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.uniform(0., 100., (n, 2))
    Y = np.array([[explore_cost_function(x[0], x[1])] for x in X]) + rng.randn(n, 1) * 0.25
    return X, Y


def log_hyp_of(gpm):
    ell = np.broadcast_to(np.asarray(gpm.kern.lengthscale, float).reshape(-1), (gpm.X.shape[1],))
    return np.log(np.r_[ell, np.sqrt(gpm.kern.variance), np.sqrt(gpm.likelihood.variance)])


def test_gp_parameter_fit_numeric_path(handle):
    import gptest_b200.gpy_compat as GPy
    X, Y = script_data()
    kernel = GPy.kern.RBF(input_dim=2, variance=10., lengthscale=20.)        # GP_parameter_fit.py:30
    gpm = GPy.models.GPRegression(X, Y, kernel)                               # :31
    lh = log_hyp_of(gpm)
    ll0 = gpm.log_likelihood()
    ref = -gpr_oracle.nlml_chol(lh, X, Y[:, 0])
    assert abs(ll0 - ref) <= 1e-8 * abs(ref)
    gpm.optimize(messages=False)                                              # :32
    ll1 = gpm.log_likelihood()
    assert ll1 > ll0 + 1.0
    g = gpr_oracle.nlml_grad(log_hyp_of(gpm), X, Y[:, 0])
    g_iso = np.r_[g[0] + g[1], g[2:]]
    assert np.abs(g_iso).max() < 1e-2 * max(1.0, abs(ll1))                    # stationary point of the isotropic model
    gpm.optimize_restarts(num_restarts=4, n_iter=25)                          # :33
    ll2 = gpm.log_likelihood()
    assert ll2 >= ll1 - 1e-9 and np.isfinite(gpm.restart_objectives).sum() >= 1
    # prediction on the script's 100 x 100 grid (:47-52); variance includes the noise (GPy semantics)
    Xt, Yt = np.meshgrid(np.arange(100), np.arange(100))
    Xfull = np.vstack([Xt.ravel(), Yt.ravel()]).transpose()
    m, v = gpm.predict(Xfull)
    assert m.shape == (10000, 1) and v.shape == (10000, 1)
    sub = np.arange(0, 10000, 97)
    rm, rv = gpr_oracle.predict_chol(log_hyp_of(gpm), X, Y[:, 0], Xfull[sub])
    assert np.abs(m[sub, 0] - rm).max() <= 1e-8 * np.abs(rm).max()
    assert np.abs(v[sub, 0] - (rv + gpm.likelihood.variance)).max() < 1e-8
    # the 40 prefix refits of the animation loop (:61-63), three of them checked
    for i in (0, 7, 39):
        gpm.set_XY(X[0:i + 1, :], Y[0:i + 1] - MEAN_VALUE)
        m, v = gpm.predict(Xfull[sub])
        rm, rv = gpr_oracle.predict_chol(log_hyp_of(gpm), X[:i + 1], Y[:i + 1, 0] - MEAN_VALUE, Xfull[sub])
        assert np.abs(m[:, 0] - rm).max() <= 1e-8 * max(np.abs(rm).max(), 1e-3)
        assert np.abs(v[:, 0] - (rv + gpm.likelihood.variance)).max() < 1e-8


def test_rbf_kernel_matrix_and_ard(handle):
    import gptest_b200.gpy_compat as GPy
    rng = np.random.default_rng(1)
    X = rng.random((90, 3))
    k = GPy.kern.RBF(3, variance=2.0, lengthscale=[0.3, 0.5, 0.9], ARD=True)
    K = k.K(X)
    from oracle import gppref_oracle
    assert np.abs(K - gppref_oracle.rbf_ard_K(X, k.lengthscale, k.variance)).max() < 1e-13
    gpm = GPy.models.GPRegression(X, np.sin(X.sum(1))[:, None], k, noise_var=0.01)
    ll = gpm.log_likelihood()
    ref = -gpr_oracle.nlml_chol(log_hyp_of(gpm), X, np.sin(X.sum(1)))
    assert abs(ll - ref) <= 1e-8 * abs(ref)
    gpm.optimize(max_iters=30)
    assert gpm.log_likelihood() > ll
    assert len(np.asarray(gpm.kern.lengthscale)) == 3


def test_replay_loop_extends_the_factor(handle, monkeypatch):
    """GP_parameter_fit.py:60-63: 40 prefixes, one more point each, a grid prediction after every set_XY.
    The factor on the device must be EXTENDED (one rebuild at the start, none after) and every frame must
    equal a refit from scratch."""
    import gptest_b200.gpy_compat as GPy
    from gptest_b200 import _lib
    X, Y = script_data()
    gpm = GPy.models.GPRegression(X, Y, GPy.kern.RBF(input_dim=2, variance=10., lengthscale=20.))
    Xt, Yt = np.meshgrid(np.arange(0, 100, 7), np.arange(0, 100, 7))
    Xgrid = np.vstack([Xt.ravel(), Yt.ravel()]).transpose().astype(float)
    h = _lib.default_handle()
    begins = []
    real_begin = h.grow_begin
    monkeypatch.setattr(h, 'grow_begin', lambda *a, **k: (begins.append(1), real_begin(*a, **k))[1])
    for ii in range(40):
        gpm.set_XY(X[0:ii + 1, :], Y[0:ii + 1] - MEAN_VALUE)
        m, v = gpm.predict(Xgrid)
        assert h.grow_size() == ii + 1
        if ii in (0, 1, 17, 39):
            rm, rv = gpr_oracle.predict_chol(log_hyp_of(gpm), X[:ii + 1], Y[:ii + 1, 0] - MEAN_VALUE, Xgrid)
            assert np.abs(m[:, 0] - rm).max() <= 1e-8 * max(np.abs(rm).max(), 1e-3)
            assert np.abs(v[:, 0] - (rv + gpm.likelihood.variance)).max() < 1e-8
    assert len(begins) == 1
    # changing a parameter invalidates the stored factor
    gpm.kern.lengthscale = np.array([25.0])
    m2, _ = gpm.predict(Xgrid)
    assert len(begins) == 2
    rm, _ = gpr_oracle.predict_chol(log_hyp_of(gpm), X[:40], Y[:40, 0] - MEAN_VALUE, Xgrid)
    assert np.abs(m2[:, 0] - rm).max() <= 1e-8 * max(np.abs(rm).max(), 1e-3)
