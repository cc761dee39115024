"""A/B of handle options on the 1024 x N=2048 sweep (BASELINE config 5): python tools/sweep_ab.py OPTION V0 V1 [B]"""
import sys
import time
import numpy as np
sys.path.insert(0, '.')
import bench_configs as cfg
from gptest_b200 import _lib

opt, v0, v1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
X, Y, lhs = cfg.make_c5(B=1024)
kh = np.array([cfg.khyp_of(l) for l in lhs[:B]])
h = _lib.Handle(0)
h.set_train(X, Y)
out = {}
for mode in (v0, v1, v0, v1):
    h.set_option(opt, mode)
    h.gpr_nlml_batched(kh[:32])
    best = 1e9
    for it in range(3):
        t0 = time.perf_counter()
        vals, info = h.gpr_nlml_batched(kh)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    out.setdefault(mode, []).append((round(best, 2), float(vals.sum())))
print(opt, out, 'max rel diff', float(np.abs(1 - np.array(out[v0][0][1]) / np.array(out[v1][0][1]))))
