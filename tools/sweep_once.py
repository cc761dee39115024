"""One batched likelihood call (config-5 shape) for launch lists: python tools/sweep_once.py [B] [n]"""
import sys
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
h.gpr_nlml_batched(kh[:8])
import time
t0 = time.perf_counter()
vals, info = h.gpr_nlml_batched(kh)
t1 = time.perf_counter()
print('B', B, 'ms', (t1 - t0) * 1e3, 'device ms', h.timings()['total_ms'], 'fail', int((info != 0).sum()), 'TF', B * n ** 3 / 3 / (t1 - t0) / 1e12)
