"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys
import numpy as np
sys.path.insert(0, '.')
from gptest_b200 import _lib
from oracle import gpr_oracle

rng = np.random.default_rng(0)
h = _lib.Handle(0)
n, d, m = 300, 3, 70
X = rng.random((n, d)); y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(n); Z = rng.random((m, d))
lh = np.log([0.5] * d + [1.0, 0.1])
kh = np.r_[np.exp(lh[:d]), np.exp(lh[d]) ** 2, np.exp(lh[d + 1]) ** 2]
h.set_train(X, y)
K = h.kxx(kh); Kxz = h.kxz(kh, Z); D2 = h.sqdist(Z)
v = h.gpr_nlml(kh); v2, g = h.gpr_nlml(kh, want_grad=True)
fz, cov = h.gpr_predict(kh, Z)
vals, info = h.gpr_nlml_batched(np.tile(kh, (5, 1)) * (1 + 0.01 * rng.random((5, d + 2))))
vals2, grads, info2 = h.gpr_nlml_batched(np.tile(kh, (3, 1)), want_grad=True)
A = rng.standard_normal((200, 200)); A = A @ A.T / 200 + np.eye(200)
L = h.potrf(A)
ref = float(gpr_oracle.nlml(lh, X, y)[0, 0])
print('nlml', v, ref, abs(v - ref) / abs(ref), 'potrf err', np.abs(L - np.linalg.cholesky(A)).max())
# Laplace paths
P = 500
uvi = rng.integers(0, n, (P, 2)); bad = uvi[:, 0] == uvi[:, 1]; uvi[bad, 1] = (uvi[bad, 0] + 1) % n
yp = np.where(rng.random(P) < 0.5, 1.0, -1.0)
h.set_train(X)
f, lml, it, tr, jit = h.pref_laplace(uvi, yp, np.r_[[0.5] * d, 1.0], max_iter=4)
W, gg = h.pref_derivatives(uvi, yp, f)
yc = np.where(rng.random(n) < 0.5, 1.0, -1.0)
f2, lml2, it2, tr2, jit2 = h.gpc_laplace(yc, np.r_[[0.5] * d, 1.0], max_iter=4)
mu, var, p = h.gpc_predict(Z)
print('laplace ok', it, it2, float(p.min()), float(p.max()))
