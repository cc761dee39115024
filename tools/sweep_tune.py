"""Config-5 sweep time vs options: python tools/sweep_tune.py"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import bench
from gptest_b200 import _lib
X, Y, lhs = bench.make_c5(2048, 1024)
kh = np.array([bench.khyp_of(l) for l in lhs])
h = _lib.Handle(0)
h.set_train(X, Y)
ref = None
for nb, chunk, split in [(2, 256, 1), (1, 256, 1), (4, 256, 1), (2, 148, 1), (2, 296, 1), (2, 512, 1), (2, 256, 0), (4, 296, 1), (4, 512, 1)]:
    h.set_option('nb_tiles', nb); h.set_option('batch_chunk', chunk); h.set_option('split_tiles', split)
    h.gpr_nlml_batched(kh[:chunk])
    t0 = time.perf_counter(); vals, info = h.gpr_nlml_batched(kh); t1 = time.perf_counter()
    if ref is None: ref = vals
    print('nb %d chunk %3d split %d: %7.2f ms  %5.2f TFLOP/s  maxrel vs first %.1e fail %d' % (nb, chunk, split, (t1 - t0) * 1e3, 1024 * 2048 ** 3 / 3 / (t1 - t0) / 1e12, np.max(np.abs(vals - ref) / np.abs(ref)), (info != 0).sum()), flush=True)
