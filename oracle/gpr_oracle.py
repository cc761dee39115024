"""CPU oracle for the GP-regression hot path (TEST INFRASTRUCTURE - not product code).

This file restates, in plain numpy, the arithmetic of the reference's ``GPr.py`` so that the
CUDA path can be checked against it.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; nothing under
``gptest_b200/`` does (the product path has no CPU fallback).

Parity status: PINNED for D == 1 - ``oracle/make_golden.py`` imports the unmodified
``/root/reference/GPr.py`` in place and checks these functions against it bit-for-bit
(KAT-1, KAT-2 in ``tests/golden/gpr_kat.json``).  For D > 1 the reference raises
(``GPr.py:5-6`` reshapes to ``(len, 1)``); the only generalisation made here is summing the
expanded-form terms over the D columns (``sqdist_expanded``), everything else is unchanged.

Every function cites the reference lines it follows.
"""
import numpy as np


def sqdist_expanded(A, B):
    """Pairwise squared distances in the reference's expanded form ``|a|^2 + |b|^2 - 2ab``.

    Follows GPr.py:4-13 (reshape, A*A, B*B, 2*dot, tile, tile, A2 + B2 - AB).  The reference
    only handles one column; here the squares and the dot product run over all D columns,
    which for D == 1 is the identical sequence of floating-point operations.
    """
    A = np.asarray(A, dtype=float)
    B = np.asarray(B, dtype=float)
    A = A.reshape(len(A), -1)                      # GPr.py:5
    B = B.reshape(len(B), -1)                      # GPr.py:6
    a2 = np.sum(A * A, axis=1, keepdims=True)      # GPr.py:7  (N,1)
    b2 = np.sum(B * B, axis=1, keepdims=True)      # GPr.py:8  (M,1)
    ab = 2 * np.dot(A, B.T)                        # GPr.py:9
    # GPr.py:10-12: the two np.tile calls only broadcast a2 / b2 to (N,M)
    return (a2 + b2.T) - ab


def split_hyp(log_hyp):
    """GPr.py:91-97: hyp = exp(logHyp); M = hyp[:n-2]; sf2 = hyp[n-2]**2; sn2 = hyp[n-1]**2."""
    hyp = np.exp(np.asarray(log_hyp, dtype=float))
    n = len(hyp)
    return hyp[:n - 2], hyp[n - 2] ** 2, hyp[n - 1] ** 2


def kxx(log_hyp, x):
    """GPr.py:99-103: sn2*I + sf2*exp(-0.5*sqdist(x/M, x/M))."""
    ell, sf2, sn2 = split_hyp(log_hyp)
    x = np.asarray(x, dtype=float)
    xs = x / ell                                   # GPr.py:100
    d2 = sqdist_expanded(xs, xs)                   # GPr.py:101
    return sn2 * np.eye(np.size(x, axis=0)) + sf2 * np.exp(-0.5 * d2)   # GPr.py:102


def kxz(log_hyp, x, z):
    """GPr.py:105-110: sf2*exp(-0.5*sqdist(x/M, z/M)) (no noise term)."""
    ell, sf2, _ = split_hyp(log_hyp)
    xs = np.asarray(x, dtype=float) / ell          # GPr.py:106
    zs = np.asarray(z, dtype=float) / ell          # GPr.py:107
    return sf2 * np.exp(-0.5 * sqdist_expanded(xs, zs))   # GPr.py:108-109


def nlml(hyp, x, y, mean=0.0):
    """Negative log marginal likelihood exactly as GPr.py:57-69 computes it.

    Explicit inverse through two general solves against the identity (GPr.py:63), quadratic
    form (GPr.py:65), sum(log(diag(L))) (GPr.py:66), n*log(2*pi)/2 (GPr.py:67).
    Returns a (1,1) ndarray like the reference.
    """
    n = np.size(x, axis=0)                         # GPr.py:58
    K = kxx(hyp, x)                                # GPr.py:59-60
    L = np.linalg.cholesky(K)                      # GPr.py:62
    iK = np.linalg.solve(L.T, np.linalg.solve(L, np.eye(n)))   # GPr.py:63
    yc = np.reshape(y, (len(y), 1))                # GPr.py:64
    err_y = np.dot(np.dot((yc - mean).T, iK), (yc - mean)) / 2   # GPr.py:65
    det_k = np.sum(np.log(np.diag(L)))             # GPr.py:66
    occam = n * np.log(2 * np.pi) / 2              # GPr.py:67
    return err_y + det_k + occam                   # GPr.py:68


def predict(log_hyp, x, y, z):
    """Posterior mean and latent variance exactly as GPr.py:45-54.

    inv(Kxx) by LU (GPr.py:48), full (M,M) product of which the diagonal is kept
    (GPr.py:50), cov = sf2 - diag (GPr.py:51,53; no sn2 added back).
    """
    Kxz = kxz(log_hyp, x, z)                       # GPr.py:46
    Kxx = kxx(log_hyp, x)                          # GPr.py:47
    iK = np.linalg.inv(Kxx)                        # GPr.py:48
    Kzx = Kxz.T                                    # GPr.py:49
    k_diag = np.diagonal(np.dot(np.dot(Kzx, iK), Kxz))   # GPr.py:50
    _, sf2, _ = split_hyp(log_hyp)
    k_noise = sf2 * np.ones(np.size(z, axis=0))    # GPr.py:51
    fz = np.dot(np.dot(Kzx, iK), np.asarray(y).T)  # GPr.py:52
    return fz, k_noise - k_diag                    # GPr.py:53-54


# ----------------------------------------------------------------------------------------
# Cholesky-based evaluation of the same quantities.  Not in the reference; used by the
# tests at sizes where inv()'s kappa*eps error would otherwise dominate the
# comparison, and as the textbook statement of what the CUDA path computes.
# ----------------------------------------------------------------------------------------
def nlml_chol(hyp, x, y):
    from scipy.linalg import cho_solve
    K = kxx(hyp, x)
    L = np.linalg.cholesky(K)
    yv = np.asarray(y, dtype=float).reshape(-1)
    alpha = cho_solve((L, True), yv)
    return 0.5 * yv @ alpha + np.sum(np.log(np.diag(L))) + len(yv) * np.log(2 * np.pi) / 2


def predict_chol(log_hyp, x, y, z):
    from scipy.linalg import cho_solve, solve_triangular
    K = kxx(log_hyp, x)
    Ks = kxz(log_hyp, x, z)
    L = np.linalg.cholesky(K)
    yv = np.asarray(y, dtype=float).reshape(-1)
    alpha = cho_solve((L, True), yv)
    V = solve_triangular(L, Ks, lower=True)
    _, sf2, _ = split_hyp(log_hyp)
    return Ks.T @ alpha, sf2 - np.sum(V * V, axis=0)


def nlml_grad(hyp, x, y):
    """Gradient of the NLML w.r.t. the log hyper-parameters [log l_1..l_D, log sf, log sn].

    The reference never computes gradients (it uses Nelder-Mead, GP_regression_demo.py:44);
    this is the textbook formula dNLML/dtheta = 0.5*tr((K^-1 - alpha alpha^T) dK/dtheta)
    (Rasmussen & Williams eq. 5.9) for the kernel of GPr.py:99-103.  Pinned by central
    differences of ``nlml`` in tests/test_oracle.py ("parity unpinned" in the reference).
    """
    from scipy.linalg import cho_solve
    x = np.asarray(x, dtype=float).reshape(len(x), -1)
    yv = np.asarray(y, dtype=float).reshape(-1)
    ell, sf2, sn2 = split_hyp(hyp)
    n, d = x.shape
    xs = x / ell
    Kse = sf2 * np.exp(-0.5 * sqdist_expanded(xs, xs))
    K = Kse + sn2 * np.eye(n)
    L = np.linalg.cholesky(K)
    alpha = cho_solve((L, True), yv)
    Q = cho_solve((L, True), np.eye(n)) - np.outer(alpha, alpha)
    g = np.empty(d + 2)
    for k in range(d):
        dk = (xs[:, k:k + 1] - xs[:, k:k + 1].T) ** 2     # d/dlog l_k of -0.5*r^2 = +r_k^2
        g[k] = 0.5 * np.sum(Q * Kse * dk)
    g[d] = 0.5 * np.sum(Q * (2.0 * Kse))
    g[d + 1] = 0.5 * np.trace(Q) * 2.0 * sn2
    return g


# ----------------------------------------------------------------------------------------
# Other stationary covariance functions behind the reference's string dispatch (GPr.py:28-32 selects by name
# and knows only "SE").  SURVEY 8f rank 4 - NOT in the reference; textbook forms (Rasmussen & Williams eq. 4.17)
# with the same hyper-parameter layout [log l_1..l_D, log sf, log sn] and the same ARD scaling as GPr.py:100.
# ----------------------------------------------------------------------------------------
KINDS = {'SE': 0, 'Matern32': 1, 'Matern52': 2}


def radial(kind, r2):
    """k(r)/sf2 and g(r) with dk/dlog l_k = sf2 * g(r) * (scaled difference in dimension k)^2."""
    r2 = np.maximum(r2, 0.0)
    if kind == 'SE':
        k = np.exp(-0.5 * r2)
        return k, k
    r = np.sqrt(r2)
    if kind == 'Matern32':
        a = np.sqrt(3.0) * r
        e = np.exp(-a)
        return (1 + a) * e, 3.0 * e
    if kind == 'Matern52':
        a = np.sqrt(5.0) * r
        e = np.exp(-a)
        return (1 + a + a * a / 3.0) * e, (5.0 / 3.0) * (1 + a) * e
    raise ValueError(kind)


def kxx_kind(log_hyp, x, kind='SE'):
    x = np.asarray(x, dtype=float).reshape(len(x), -1)
    ell, sf2, sn2 = split_hyp(log_hyp)
    xs = x / ell
    return sf2 * radial(kind, sqdist_expanded(xs, xs))[0] + sn2 * np.eye(len(x))


def kxz_kind(log_hyp, x, z, kind='SE'):
    x = np.asarray(x, dtype=float).reshape(len(x), -1)
    z = np.asarray(z, dtype=float).reshape(len(z), -1)
    ell, sf2, _ = split_hyp(log_hyp)
    return sf2 * radial(kind, sqdist_expanded(x / ell, z / ell))[0]


def nlml_kind(log_hyp, x, y, kind='SE', want_grad=False):
    from scipy.linalg import cho_solve
    x = np.asarray(x, dtype=float).reshape(len(x), -1)
    yv = np.asarray(y, dtype=float).reshape(-1)
    ell, sf2, sn2 = split_hyp(log_hyp)
    n, d = x.shape
    xs = x / ell
    kr, gr = radial(kind, sqdist_expanded(xs, xs))
    K = sf2 * kr + sn2 * np.eye(n)
    L = np.linalg.cholesky(K)
    alpha = cho_solve((L, True), yv)
    val = 0.5 * yv @ alpha + np.sum(np.log(np.diag(L))) + n * np.log(2 * np.pi) / 2
    if not want_grad:
        return val
    Q = cho_solve((L, True), np.eye(n)) - np.outer(alpha, alpha)
    g = np.empty(d + 2)
    for k in range(d):
        g[k] = 0.5 * np.sum(Q * sf2 * gr * (xs[:, k:k + 1] - xs[:, k:k + 1].T) ** 2)
    g[d] = 0.5 * np.sum(Q * 2.0 * sf2 * kr)
    g[d + 1] = 0.5 * np.trace(Q) * 2.0 * sn2
    return val, g


def predict_kind(log_hyp, x, y, z, kind='SE'):
    from scipy.linalg import cho_solve, solve_triangular
    K = kxx_kind(log_hyp, x, kind)
    Ks = kxz_kind(log_hyp, x, z, kind)
    L = np.linalg.cholesky(K)
    alpha = cho_solve((L, True), np.asarray(y, dtype=float).reshape(-1))
    V = solve_triangular(L, Ks, lower=True)
    return Ks.T @ alpha, split_hyp(log_hyp)[1] - np.sum(V * V, axis=0)
