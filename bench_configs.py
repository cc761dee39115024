"""Synthetic inputs of the BASELINE.json configs (SURVEY section 8d recipes), shared by bench.py and the
full-size parity tests so that both run exactly the same data.  Everything is numpy, seeded, fp64."""
from math import erfc, sqrt

import numpy as np

N_FIT, D_FIT, M_TEST = 16384, 8, 1024


def _ndtr(v):
    return np.vectorize(lambda t: 0.5 * erfc(-t / sqrt(2.0)))(v)


def khyp_of(log_hyp):
    """[l_1..l_D, sf2, sn2] from log hyper-parameters [log l.., log sf, log sn] (GPr.py:93-97)."""
    h = np.exp(np.asarray(log_hyp, dtype=float))
    return np.concatenate([h[:-2], [h[-2] ** 2, h[-1] ** 2]])


def make_c2(n=N_FIT, d=D_FIT, m=M_TEST):
    """C2: GPr regression N=16384, D=8 SE-ARD; well conditioned (cond ~ 2e5)."""
    rng = np.random.default_rng(0)
    X = rng.random((n, d))
    w = rng.standard_normal(d)
    y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
    Z = rng.random((m, d))
    log_hyp = np.log([0.5] * d + [1.0, 0.1])
    return X, y, Z, log_hyp


def make_c3(n=8192, d=4, m=1024):
    """C3: GPc labels from Bernoulli(Phi(latent)) (GP_classification_demo.py:10-21, GPc.py:37)."""
    rng = np.random.default_rng(0)
    X = rng.random((n, d))
    w = rng.standard_normal(d)
    lat = np.sin(2 * np.pi * X @ w / np.abs(w).sum() + np.pi / 4) + 0.2
    y = np.where(rng.random(n) < _ndtr(lat), 1.0, -1.0)
    Z = rng.random((m, d))
    loghyp = np.log([0.5] * d + [1.0])
    return X, y, Z, loghyp


def make_c4(n=4096, d=6, P=32768):
    """C4: GPpref items / pairs / noisy rank labels (GP_preference_demo.py:11,20-29); loghyp layout GPpref.py:99."""
    rng = np.random.default_rng(0)
    X = rng.random((n, d))
    uvi = rng.integers(0, n, (P, 2))
    bad = uvi[:, 0] == uvi[:, 1]
    uvi[bad, 1] = (uvi[bad, 0] + 1) % n
    w = rng.standard_normal(d)
    lat = np.sin(2 * np.pi * X @ w / np.abs(w).sum() + np.pi / 4) + 0.2
    fu = lat[uvi[:, 0]] + 0.05 * rng.standard_normal(P)
    fv = lat[uvi[:, 1]] + 0.05 * rng.standard_normal(P)
    y = np.where(fv > fu, 1.0, -1.0).reshape(-1, 1)
    loghyp = np.log([0.5] * d + [1.0, 0.1])
    return X, uvi, y, loghyp


def make_c5(n=2048, B=1024):
    """C5: GP_parameter_fit.py:9-28 data, 32x32 grid over (log l, log sf), sn = 0.25."""
    rng = np.random.default_rng(0)
    X = 100 * rng.random((n, 2))
    a, b = X[:, 0], X[:, 1]
    cost = 3.0 + 10 * np.exp(-np.sqrt((a - 40) ** 2 + (b - 40) ** 2) / 16) \
        + 7 * np.exp(-np.sqrt((a - 10) ** 2 + (b - 90) ** 2) / 12) \
        + 4 * np.exp(-np.sqrt((a - 80) ** 2 + (b - 60) ** 2) / 32) \
        + 7 * np.exp(-np.sqrt((a + 20) ** 2 + (b - 50) ** 2) / 32) \
        + 7 * np.exp(-np.sqrt((a - 120) ** 2 + (b - 50) ** 2) / 32) \
        + 12 * np.exp(-np.sqrt((a - 80) ** 2 + (b - 20) ** 2) / 8) \
        + 5 * np.exp(-np.sqrt((a - 60) ** 2 + (b - 80) ** 2) / 10) \
        + 3 * np.exp(-np.sqrt((a - 90) ** 2 + (b - 90) ** 2) / 20)
    Y = cost + 0.25 * rng.standard_normal(n) - 3.0
    g = int(round(np.sqrt(B)))
    ll = np.linspace(np.log(2), np.log(200), g)
    lf = np.linspace(np.log(0.3), np.log(30), g)
    L, F = np.meshgrid(ll, lf, indexing='ij')
    lh = np.stack([L.ravel(), L.ravel(), F.ravel(), np.full(g * g, np.log(0.25))], axis=1)[:B]
    return X, Y, lh
