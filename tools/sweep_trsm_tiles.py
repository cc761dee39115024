"""Panel TRSM CTA-tile switch (option trsm_tile_threshold) on the C5 sweep: python tools/sweep_trsm_tiles.py [B]"""
import sys
import numpy as np
sys.path.insert(0, '.')
import bench_configs as cfg
from gptest_b200 import _lib, sweep
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
X, Y, lhs = cfg.make_c5()
h = _lib.Handle(0)
h.set_train(X, Y)
kh = sweep.natural_params(lhs[:B])
ref = None
for thr in (2400, 1 << 40, 8000, 2400, 1 << 40, 8000):
    h.set_option('trsm_tile_threshold', thr)
    h.gpr_nlml_batched(kh)
    ts = []
    for i in range(3):
        vals, info = h.gpr_nlml_batched(kh)
        ts.append(h.timings()['total_ms'])
    if ref is None:
        ref = vals
    print('B', B, 'trsm_tile_threshold', thr, 'ms', round(min(ts), 3), 'bitwise equal:', bool(np.array_equal(ref, vals)), flush=True)
