"""Batched value + gradient (the multi-start L-BFGS evaluation): python tools/sweep_grad.py [B] [n]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import bench
from gptest_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
X, Y, lhs = bench.make_c5(n, 1024)
kh = np.array([bench.khyp_of(l) for l in lhs[:B]])
h = _lib.Handle(0)
h.set_train(X, Y)
h.gpr_nlml_batched(kh, want_grad=True)
t0 = time.perf_counter()
vals, grads, info = h.gpr_nlml_batched(kh, want_grad=True)
dt = time.perf_counter() - t0
h.gpr_nlml_batched(kh)
t0 = time.perf_counter()
h.gpr_nlml_batched(kh)
dv = time.perf_counter() - t0
print('B', B, 'n', n, 'value+grad ms %.2f (%.2f TF over N^3)' % (dt * 1e3, B * n ** 3 / dt / 1e12), 'value only ms %.2f' % (dv * 1e3),
      'ratio %.2f' % (dt / dv), 'timings', h.timings())
