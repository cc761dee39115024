"""Per-call overhead of the batched path at the 8-GPU slice size (128 problems): python tools/sweep_small.py"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import bench
from gptest_b200 import _lib
X, Y, lhs = bench.make_c5(2048, 1024)
kh = np.array([bench.khyp_of(l) for l in lhs])
h = _lib.Handle(0)
h.set_train(X, Y)
for B in (128, 256, 512, 1024):
    h.gpr_nlml_batched(kh[:B])
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); h.gpr_nlml_batched(kh[:B]); ts.append((time.perf_counter() - t0) * 1e3)
    print('B %4d: %s ms  device %.2f ms -> %.2f TFLOP/s' % (B, ['%.2f' % t for t in ts], h.timings()['total_ms'], B * 2048 ** 3 / 3 / min(ts) / 1e9))
