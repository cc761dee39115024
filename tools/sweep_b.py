"""One batched likelihood call for B problems of the C5 recipe (ncu launch-list target): python tools/sweep_b.py [B]"""
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
h.gpr_nlml_batched(kh)
vals, info = h.gpr_nlml_batched(kh)
print(vals[:3], h.timings())
