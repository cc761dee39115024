"""CPU oracle for the preference-GP Laplace path (TEST INFRASTRUCTURE - not product code).

Python-3 / numpy restatement of the reference's ``GPpref.py`` (which is Python 2 and imports
the un-vendored GPy, so it cannot run here as-is).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this file.

Parity status
  * Laplace loop, likelihood derivatives, log_marginal: PINNED - ``oracle/make_golden.py``
    executes the reference's own ``GPpref.py`` source (its two py2 ``print`` statements
    rewritten in memory, nothing else) and checks this restatement against it on the demo
    data and on data with repeated items (KAT-3, KAT-4 in ``tests/golden/gppref_kat.npz``).
  * Covariance matrix: PARITY UNPINNED - the reference obtains K from
    ``GPy.kern.RBF(ARD=True).K`` (GPpref.py:109,122); GPy is a third-party dependency that
    is absent from /root/reference and pinned nowhere (no requirements/lock file).
    ``rbf_ard_K`` restates GPy's published RBF/Stationary algorithm from its documentation.

The reference's quirks are reproduced on purpose:
  1. gradient scatter with fancy-index ``+=`` : last write wins on repeated items
     (GPpref.py:77-78), while W accumulates (GPpref.py:82-87);
  2. ``calc_laplace`` overwrites the *method* ``set_sigma`` (GPpref.py:115) so the probit
     sigma stays at its constructor value 1.0;
  3. ``logdetK = sum(log(diag(L)))`` is half the log-determinant (GPpref.py:131) and is
     halved again in ``log_marginal`` (GPpref.py:93).
"""
import numpy as np
from scipy.special import ndtr

_SQRT_2PI = np.sqrt(2 * np.pi)


def std_norm_pdf(x):
    """GPpref.py:7-10."""
    x = np.clip(x, -1e150, 1e150)
    return np.exp(-(x ** 2) / 2) / _SQRT_2PI


def rbf_ard_K(x, lengthscale, variance):
    """GPy ``RBF(ARD=True).K(X)`` restated (GPy Stationary._scaled_dist / RBF.K_of_r).

    X is divided by the lengthscales; r2 = -2 X X^T + |X|^2 + |X|^2^T with the diagonal forced
    to zero and negatives clipped to zero; r = sqrt(r2); K = variance * exp(-0.5 * r**2).
    """
    xs = np.asarray(x, dtype=float) / np.asarray(lengthscale, dtype=float)
    xsq = np.sum(np.square(xs), axis=1)
    r2 = -2.0 * np.dot(xs, xs.T) + (xsq[:, None] + xsq[None, :])
    np.fill_diagonal(r2, 0.0)
    r2 = np.clip(r2, 0, np.inf)
    r = np.sqrt(r2)
    return variance * np.exp(-0.5 * r ** 2)


class ProbitPrefOracle:
    """Restatement of ``PrefProbit`` (GPpref.py:46-94)."""

    def __init__(self, sigma=1.0):
        self.sigma = sigma                                   # GPpref.py:51-54
        self.isqrt2sig = 1.0 / (sigma * np.sqrt(2.0))
        self.i2var = self.isqrt2sig ** 2
        self.log2pi = np.log(2.0 * np.pi)                    # GPpref.py:49

    def z_k(self, uvi, f, y):
        """GPpref.py:56-58: y * (f[v] - f[u]) / (sqrt(2) sigma); shapes (P,1)."""
        return y * (self.isqrt2sig * (f[uvi[:, 1]] - f[uvi[:, 0]]))

    def derivatives(self, uvi, y, f, accumulate=False):
        """GPpref.py:68-88 including the last-write-wins gradient (quirk 1).

        accumulate=True (NOT the reference) sums the contributions of repeated items instead,
        which makes the loop a true Newton iteration; used to test the product's opt-in mode."""
        nx = len(f)
        z = self.z_k(uvi, f, y)
        phi_z = ndtr(z)                                      # GPpref.py:71
        n_z = std_norm_pdf(z)                                # GPpref.py:72
        g = np.zeros((nx, 1), dtype=float)                   # GPpref.py:75
        d = y * self.isqrt2sig * n_z / phi_z                 # GPpref.py:76
        # GPpref.py:77-78.  ``a[idx] += v`` is ``a[idx] = a[idx] + v``: one gather, one add,
        # one scatter; for a repeated index the LAST occurrence is the value that stays.
        if accumulate:
            np.add.at(g, (uvi[:, 0], 0), -d[:, 0])
            np.add.at(g, (uvi[:, 1], 0), d[:, 0])
        else:
            g[uvi[:, 0]] = g[uvi[:, 0]] - d
            g[uvi[:, 1]] = g[uvi[:, 1]] + d
        inner = -self.i2var * (z * n_z / phi_z + (n_z / phi_z) ** 2)   # GPpref.py:80
        W = np.zeros((nx, nx), dtype=float)                  # GPpref.py:81
        w = -inner[:, 0]                                     # W[..] -= ddpy_df  ==  += w
        u, v = uvi[:, 0], uvi[:, 1]
        # GPpref.py:82-87: sequential accumulation, pair by pair.  np.add.at is unbuffered
        # and visits the operands in order, so every cell sees the same sequence of adds.
        rows = np.stack([u, v, u, v], axis=1).ravel()        # per pair: (u,u) (v,v) (u,v) (v,u)
        cols = np.stack([u, v, v, u], axis=1).ravel()
        vals = np.stack([w, w, -w, -w], axis=1).ravel()
        np.add.at(W, (rows, cols), vals)
        return W, g

    def log_marginal(self, uvi, y, f, iK, logdetK):
        """GPpref.py:90-94 (quirk 3 lives in the caller's logdetK)."""
        z = self.z_k(uvi, f, y)
        phi_z = ndtr(z)
        psi = (np.sum(np.log(phi_z)) - 0.5 * np.matmul(np.matmul(f.T, iK), f)
               - 0.5 * logdetK - iK.shape[0] / 2.0 * self.log2pi)
        return psi.flat[0]


def gradient_last_writer(uvi, n):
    """Index form of quirk 1, shared with the tests of the device kernel.

    Returns (ku, kv): for every item i the index of the LAST pair whose u (resp. v) is i, or
    -1.  The reference gradient is  g[i] = -d[ku[i]] (if any) + d[kv[i]] (if any).
    """
    ku = np.full(n, -1, dtype=np.int64)
    kv = np.full(n, -1, dtype=np.int64)
    ku[uvi[:, 0]] = np.arange(len(uvi))       # later assignments overwrite earlier ones
    kv[uvi[:, 1]] = np.arange(len(uvi))
    return ku, kv


def calc_laplace(x, uvi, y, loghyp, delta_f=1e-6, f=None, max_iter=None, return_trace=False, accumulate=False):
    """``PreferenceGaussianProcess.calc_laplace`` (GPpref.py:112-157).

    ``max_iter`` (not in the reference, which has no cap) bounds test/bench runs.
    Returns (f (n,1), lml) and, with return_trace, the per-iteration (f_error, lml) list the
    reference prints (GPpref.py:154).
    """
    x = np.asarray(x, dtype=float)
    n, d = x.shape
    y = np.asarray(y, dtype=float).reshape(-1, 1)
    lik = ProbitPrefOracle()                                 # sigma = 1.0, never updated (quirk 2)
    lengthscale = np.exp(loghyp[0:d])                        # GPpref.py:113
    variance = (np.exp(loghyp[d])) ** 2                      # GPpref.py:114
    if f is None:
        f = np.zeros((n, 1))                                 # GPpref.py:117-118
    Ix = np.eye(n)                                           # GPpref.py:121
    K = rbf_ard_K(x, lengthscale, variance)                  # GPpref.py:122
    eps = 1e-6                                               # GPpref.py:123
    while True:                                              # GPpref.py:126-135
        try:
            L = np.linalg.cholesky(K + eps * Ix)
            iK = np.linalg.solve(L.T, np.linalg.solve(L, Ix))
            logdetK = np.sum(np.log(L.diagonal()))           # = 0.5*log|K| (quirk 3)
            break
        except np.linalg.LinAlgError:
            eps = eps * 10
    f_error = delta_f + 1                                    # GPpref.py:138
    trace = []
    lml = None
    while f_error > delta_f:                                 # GPpref.py:140
        W, g = lik.derivatives(uvi, y, f, accumulate)        # GPpref.py:141
        G = iK + W                                           # GPpref.py:142
        f_new = np.matmul(np.linalg.inv(G), np.matmul(W, f) + g)   # GPpref.py:143
        lml = lik.log_marginal(uvi, y, f_new, iK, logdetK)   # GPpref.py:144
        f_error = np.max(np.abs(f_new - f))                  # GPpref.py:151-152
        trace.append((float(f_error), float(lml)))
        f = f_new                                            # GPpref.py:155
        if max_iter is not None and len(trace) >= max_iter:
            break
    if return_trace:
        return f, lml, trace
    return f, lml


# ---------------------------------------------------------------------------------------------------------
# Opt-in extensions.  NOT in the reference: its log_marginal (GPpref.py:90-94) has no W term
# and it has no prediction at all.  Restated from Rasmussen & Williams ch. 3 for the checker of
# gpb_pref_evidence / gpb_pref_predict; parity is unpinned by construction.
# ---------------------------------------------------------------------------------------------------------
def _k_and_w(x, uvi, y, loghyp, f, sigma, eps):
    x = np.asarray(x, dtype=float).reshape(len(x), -1)
    d = x.shape[1]
    ell = np.exp(loghyp[:d])
    sf2 = np.exp(loghyp[d]) ** 2
    K = rbf_ard_K(x, ell, sf2) + eps * np.eye(len(x))          # the jittered matrix the fit used (GPpref.py:123-128)
    lik = ProbitPrefOracle(sigma)
    f = np.asarray(f, dtype=float).reshape(-1, 1)
    W, _ = lik.derivatives(np.asarray(uvi), np.asarray(y, dtype=float).reshape(-1, 1), f, accumulate=True)
    return x, ell, sf2, K, W, lik, f


def laplace_evidence(x, uvi, y, loghyp, f, sigma=1.0, eps=1e-6):
    """R&W eq. 3.32 at the mode f: sum log Phi(z) - f' K^-1 f / 2 - log|I + K W| / 2."""
    x, ell, sf2, K, W, lik, f = _k_and_w(x, uvi, y, loghyp, f, sigma, eps)
    z = lik.z_k(np.asarray(uvi), f, np.asarray(y, dtype=float).reshape(-1, 1))
    iK = np.linalg.inv(K)
    sign, logdet = np.linalg.slogdet(np.eye(len(x)) + K @ W)
    return float(np.sum(np.log(ndtr(z))) - 0.5 * (f.T @ iK @ f).flat[0] - 0.5 * logdet)


def predict_latent(x, uvi, y, loghyp, f, z, zb=None, sigma=1.0, eps=1e-6):
    """Latent posterior at test items z (or of the difference f(zb) - f(z)): mean k*' K^-1 f,
    var k** - k*' (K + W^-1)^-1 k* written without W^-1 (W is singular):
    k** - k*' K^-1 k* + (K^-1 k*)' (K^-1 + W)^-1 (K^-1 k*).  With zb also Phi(mean / sqrt(2 sigma^2 + var))."""
    x, ell, sf2, K, W, lik, f = _k_and_w(x, uvi, y, loghyp, f, sigma, eps)

    def cross(zz):
        zz = np.asarray(zz, dtype=float).reshape(len(zz), -1)
        xs, zs = x / ell, zz / ell
        r2 = np.clip(np.sum(zs * zs, 1)[:, None] + np.sum(xs * xs, 1)[None, :] - 2 * zs @ xs.T, 0, np.inf)
        return sf2 * np.exp(-0.5 * r2), zs                       # (m, n)

    R, zs = cross(z)
    kss = np.full(len(R), sf2)
    if zb is not None:
        Rb, zbs = cross(zb)
        R = Rb - R
        kss = 2 * sf2 - 2 * sf2 * np.exp(-0.5 * np.sum((zs - zbs) ** 2, axis=1))
    iK = np.linalg.inv(K)
    T = R @ iK
    mean = (T @ f)[:, 0]
    G = iK + W
    var = kss - np.sum(R * T, axis=1) + np.sum(T * np.linalg.solve(G, T.T).T, axis=1)
    if zb is None:
        return mean, var
    return mean, var, ndtr(mean / np.sqrt(2 * sigma ** 2 + np.maximum(var, 0)))
