"""K-build time inside a fit (lower tiles, mode 1) at N=16384 and inside a 256-problem sweep chunk at N=2048."""
import sys
sys.path.insert(0, '.')
import numpy as np
import bench_configs as cfg
from gptest_b200 import _lib
h = _lib.Handle(0)
X, y, Z, lh = cfg.make_c2()
kh = cfg.khyp_of(lh)
h.set_train(X, y)
h.gpr_nlml(kh)
ts = []
for i in range(6):
    h.gpr_nlml(kh)
    ts.append(h.timings()['kbuild_ms'])
n = X.shape[0]
byt = 4.0 * n * (n + 1) + 8.0 * n * X.shape[1]
best = min(ts)
print('kbuild_ms', [round(t, 4) for t in ts], 'best', round(best, 4), 'GB/s', round(byt / best / 1e6, 1), flush=True)
