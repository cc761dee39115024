"""A few Newton iterations of the preference Laplace path at C4 size for ncu launch lists:
   python tools/pref_once.py [n] [P] [iters]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from gptest_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
P = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
D = 6
rng = np.random.default_rng(0)
X = rng.random((n, D))
uvi = rng.integers(0, n, (P, 2))
bad = uvi[:, 0] == uvi[:, 1]
uvi[bad, 1] = (uvi[bad, 0] + 1) % n
w = rng.standard_normal(D)
lat = np.sin(2 * np.pi * X @ w / np.abs(w).sum() + np.pi / 4) + 0.2
y = np.where(lat[uvi[:, 1]] + 0.05 * rng.standard_normal(P) > lat[uvi[:, 0]] + 0.05 * rng.standard_normal(P), 1.0, -1.0)
h = _lib.Handle(0)
h.set_train(X)
f, lml, it, trace, jit = h.pref_laplace(uvi, y, np.r_[[0.5] * D, 1.0], delta_f=1e-6, max_iter=iters)
print(it, lml, h.timings())
