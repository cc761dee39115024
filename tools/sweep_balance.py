"""Panel TRSM warp balance (option trsm_balance 0/1/2) on a B-problem slice of the C5 sweep: python tools/sweep_balance.py [B]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import bench_configs as cfg
from gptest_b200 import _lib, sweep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
X, Y, lhs = cfg.make_c5()
h = _lib.Handle(0)
h.set_train(X, Y)
kh = sweep.natural_params(lhs[:B])
ref = None
for mode in (0, 1, 0, 1):
    h.set_option('trsm_balance', mode)
    h.gpr_nlml_batched(kh)
    ts = []
    for i in range(5):
        vals, info = h.gpr_nlml_batched(kh)
        ts.append(h.timings()['total_ms'])
    if ref is None:
        ref = vals
    print('B', B, 'trsm_balance', mode, 'ms', round(min(ts), 3), 'bitwise equal to mode 0:', bool(np.array_equal(ref, vals)), flush=True)
