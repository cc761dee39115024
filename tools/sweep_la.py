"""Batched sweep with / without the look-ahead schedule: python tools/sweep_la.py B n [batch_chunk]"""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
import bench
from gptest_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
X, Y, lhs = bench.make_c5(n, 1024)
kh = np.array([bench.khyp_of(l) for l in lhs[:B]])
h = _lib.Handle(0)
h.set_train(X, Y)
if len(sys.argv) > 3:
    h.set_option('batch_chunk', int(sys.argv[3]))
for la in (1, 100000, 1):
    h.set_option('la_max_batch', la)
    h.gpr_nlml_batched(kh)
    t0 = time.perf_counter()
    for _ in range(3):
        vals, info = h.gpr_nlml_batched(kh)
    t1 = time.perf_counter()
    print('la_max_batch', la, 'B', B, 'n', n, 'ms', (t1 - t0) / 3 * 1e3, 'TF', B * n ** 3 / 3 / ((t1 - t0) / 3) / 1e12)
