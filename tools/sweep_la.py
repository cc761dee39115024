"""python tools/sweep_la.py B n la_max_batch [la_max_batch ...]"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import bench
from gptest_b200 import _lib
B, n = int(sys.argv[1]), int(sys.argv[2])
X, Y, lhs = bench.make_c5(n, 1024)
kh = np.array([bench.khyp_of(l) for l in lhs[:B]])
h = _lib.Handle(0)
h.set_train(X, Y)
for la in [int(a) for a in sys.argv[3:]] * 2:
    h.set_option('la_max_batch', la)
    h.gpr_nlml_batched(kh)
    t0 = time.perf_counter()
    for _ in range(3):
        vals, info = h.gpr_nlml_batched(kh)
    dt = (time.perf_counter() - t0) / 3
    print('B', B, 'n', n, 'la_max_batch', la, 'ms %.3f' % (dt * 1e3), 'TF %.2f' % (B * n ** 3 / 3 / dt / 1e12), flush=True)
