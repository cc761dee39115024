"""Batched-sweep time under settings of one or two options: python tools/sweep_opt.py B opt=v1,v2 [opt2=w1,w2]"""
import itertools
import sys
import numpy as np
sys.path.insert(0, '.')
import bench_configs as cfg
from gptest_b200 import _lib, sweep
B = int(sys.argv[1])
opts = [(a.split('=')[0], [int(v) for v in a.split('=')[1].split(',')]) for a in sys.argv[2:]]
X, Y, lhs = cfg.make_c5()
h = _lib.Handle(0)
h.set_train(X, Y)
kh = sweep.natural_params(lhs[:B])
ref = None
for rep in range(2):
    for combo in itertools.product(*[v for _, v in opts]):
        for (name, _), v in zip(opts, combo):
            h.set_option(name, v)
        h.gpr_nlml_batched(kh)
        ts = []
        for i in range(3):
            vals, info = h.gpr_nlml_batched(kh)
            ts.append(h.timings()['total_ms'])
        if ref is None:
            ref = vals
        print('B', B, dict(zip([n for n, _ in opts], combo)), 'ms', round(min(ts), 3), 'same bits:', bool(np.array_equal(ref, vals)), flush=True)
