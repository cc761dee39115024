"""Host overhead of the product API on the 8-GPU slice size: raw batched call vs sweep.sweep_nlml for B problems (1 rank)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import bench_configs as cfg
from gptest_b200 import _lib, sweep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
X, Y, lhs = cfg.make_c5()
lhs = lhs[:B]
h = _lib.default_handle()
h.set_train(X, Y)
kh = sweep.natural_params(lhs)
for name, fn in (('raw gpr_nlml_batched', lambda: h.gpr_nlml_batched(kh)), ('sweep_nlml (1 rank)', lambda: sweep.sweep_nlml(X, Y, lhs)),
                 ('set_train only', lambda: h.set_train(X, Y))):
    fn(); fn()
    ts = []
    for i in range(5):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    print('%-24s B=%d  best %.3f ms  median %.3f ms' % (name, B, min(ts), sorted(ts)[2]), h.timings()['total_ms'])
