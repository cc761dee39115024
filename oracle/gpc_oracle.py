"""CPU oracle for binary GP classification by Laplace approximation (TEST INFRASTRUCTURE).

PARITY UNPINNED: the reference's ``GPc.py`` is an unfinished fragment that does not parse
(SyntaxError at GPc.py:10) and ``GP_classification_demo.py`` never fits a model, so there is
no reference arithmetic to run.  This oracle is the specification, written from what the
fragment does state and what it cites:

  * likelihood ``log p(y|f) = log Phi(y f)`` (GPc.py:5-6), probit by default, logistic when
    ``inverse_link_function == 'Logit'`` (GPc.py:13-21);
  * labels in {0,1} are mapped to {-1,+1}, anything else is an AssertionError (GPc.py:24-38);
  * ``Psi(f) = log p(y|f) + log p(f|X)``, Rasmussen & Williams eq. 3.12 (GPc.py:42-44):
    Newton mode finding = R&W Algorithm 3.1, prediction = Algorithm 3.2,
    class probability ``Phi(mu / sqrt(1 + var))`` (R&W eq. 3.82 for the probit link);
  * loop conventions borrowed from the sibling ``GPpref.py``: f starts at 0 (GPpref.py:117-118),
    K gets a jitter ``eps*I`` starting at 1e-6 and multiplied by 10 on LinAlgError
    (GPpref.py:123-135), iteration continues while ``max|f_new - f| > delta_f``
    (GPpref.py:138-152), covariance is SE-ARD with loghyp = [l_1..l_D, sigma_f]
    (GPpref.py:99,113-114 without the probit-sigma slot).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this file.
"""
import numpy as np
from scipy.special import ndtr, log_ndtr

from .gppref_oracle import rbf_ard_K, std_norm_pdf


def preprocess_labels(y):
    """GPc.py:24-38."""
    y = np.asarray(y).astype(float).reshape(-1)
    n1 = np.sum(y == 1)
    n0 = np.sum(y == 0)
    nm = np.sum(y == -1)
    assert (n1 + n0 == y.size) or (n1 + nm == y.size), 'Inputs should be in {0,1} or {-1,1}.'
    y = y.copy()
    y[y == 0] = -1
    return y


def probit_terms(y, f):
    """log p, d log p / df, W = -d2 log p / df2 for p = Phi(y f) (R&W eq. 3.16)."""
    yf = y * f
    lp = log_ndtr(yf)
    r = np.exp(-0.5 * f * f - 0.5 * np.log(2 * np.pi) - lp)      # N(f)/Phi(yf), stable
    g = y * r
    W = r * r + yf * r
    return lp, g, W


def logit_terms(y, f):
    """Same for the logistic link (GPc.py:13-14), R&W eq. 3.15."""
    yf = y * f
    lp = -np.logaddexp(0.0, -yf)
    pi = 1.0 / (1.0 + np.exp(-f))
    g = (y + 1) / 2 - pi
    W = pi * (1 - pi)
    return lp, g, W


def calc_laplace(x, y, loghyp, link='probit', delta_f=1e-6, f=None, max_iter=100,
                 return_state=False):
    """R&W Algorithm 3.1.  Returns (f (n,), lml) [, state dict for prediction]."""
    x = np.asarray(x, dtype=float).reshape(len(x), -1)
    y = preprocess_labels(y)
    n, d = x.shape
    terms = logit_terms if link == 'Logit' else probit_terms
    K0 = rbf_ard_K(x, np.exp(loghyp[:d]), np.exp(loghyp[d]) ** 2)
    eps = 1e-6
    while True:
        try:
            K = K0 + eps * np.eye(n)
            np.linalg.cholesky(K)
            break
        except np.linalg.LinAlgError:
            eps *= 10
    f = np.zeros(n) if f is None else np.asarray(f, dtype=float).reshape(-1).copy()
    f_error = delta_f + 1
    it = 0
    trace = []
    a = np.zeros(n)
    while f_error > delta_f and it < max_iter:
        lp, g, W = terms(y, f)
        sW = np.sqrt(W)
        B = np.eye(n) + sW[:, None] * K * sW[None, :]            # Alg 3.1 line 5
        L = np.linalg.cholesky(B)
        b = W * f + g                                            # line 6
        t = np.linalg.solve(L.T, np.linalg.solve(L, sW * (K @ b)))
        a = b - sW * t                                           # line 7
        f_new = K @ a                                            # line 8
        f_error = np.max(np.abs(f_new - f))
        f = f_new
        it += 1
        lp_new = terms(y, f)[0]
        trace.append((float(f_error), float(-0.5 * a @ f + np.sum(lp_new))))
    # Approximate log marginal likelihood at the returned f (Alg 3.1 line 10): W and L are
    # re-evaluated at the final f; a is the last iteration's (f = K a exactly, line 8).
    lp, g, W = terms(y, f)
    sW = np.sqrt(W)
    L = np.linalg.cholesky(np.eye(n) + sW[:, None] * K * sW[None, :])
    lml = -0.5 * a @ f + np.sum(lp) - np.sum(np.log(np.diag(L)))
    if return_state:
        return f, lml, dict(K=K, L=L, sW=sW, g=g, it=it, trace=trace, eps=eps)
    return f, lml


def predict(x, y, loghyp, z, link='probit', delta_f=1e-6, max_iter=100):
    """R&W Algorithm 3.2: latent mean, latent variance, class probability at z."""
    x = np.asarray(x, dtype=float).reshape(len(x), -1)
    z = np.asarray(z, dtype=float).reshape(len(z), -1)
    d = x.shape[1]
    f, lml, st = calc_laplace(x, y, loghyp, link, delta_f, max_iter=max_iter, return_state=True)
    ell = np.exp(loghyp[:d])
    sf2 = np.exp(loghyp[d]) ** 2
    xs, zs = x / ell, z / ell
    r2 = np.clip(np.sum(xs * xs, 1)[:, None] + np.sum(zs * zs, 1)[None, :] - 2 * xs @ zs.T, 0, np.inf)
    Ks = sf2 * np.exp(-0.5 * r2)                                 # (n, m)
    mu = Ks.T @ st['g']                                          # Alg 3.2 line 4
    v = np.linalg.solve(st['L'], st['sW'][:, None] * Ks)         # line 5
    var = sf2 - np.sum(v * v, axis=0)                            # line 6 (k** = sf2, latent)
    if link == 'Logit':
        p = 1.0 / (1.0 + np.exp(-mu / np.sqrt(1 + np.pi * var / 8)))   # MacKay's approximation
    else:
        p = ndtr(mu / np.sqrt(1 + var))                          # R&W eq. 3.82
    return mu, var, p
