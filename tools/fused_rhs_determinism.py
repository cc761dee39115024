import sys
sys.path.insert(0, '.')
import numpy as np
import bench_configs as cfg
from gptest_b200 import _lib
h = _lib.Handle(0)
h.set_option('fuse_min_tiles', 0)      # force the fused solve at every size
for n in (4096, 9216, 12288, 16384):
    X, y, Z, lh = cfg.make_c2(n=n)
    kh = cfg.khyp_of(lh)
    h.set_train(X, y)
    h.set_option('fuse_rhs', 0)
    ref = h.gpr_nlml(kh)
    h.set_option('fuse_rhs', 1)
    vals = []
    for i in range(12):
        try:
            vals.append(h.gpr_nlml(kh))
        except Exception as e:
            vals.append('LinAlgError')
    print(n, len(set(vals)), 'distinct', list(set(vals))[:3], 'unfused', ref, flush=True)
import time
X, y, Z, lh = cfg.make_c2()
kh = cfg.khyp_of(lh); h.set_train(X, y)
for f in (0, 1, 0, 1):
    h.set_option('fuse_rhs', f)
    h.gpr_nlml(kh)
    ts = []
    for i in range(5):
        h.gpr_nlml(kh); ts.append(round(h.timings()['factor_ms'], 3))
    print('fuse', f, 'factor ms', ts, flush=True)
