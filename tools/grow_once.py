"""Appending to a stored factor vs refitting (GP_parameter_fit.py:61-63 at scale):
    python tools/grow_once.py [N] [D] [step] [n_appends] [grid]
Starts from N - step*n_appends points, appends `step` points n_appends times (timing each), predicts `grid`
points from the stored factor, and times the full refit + predict of the final set for comparison."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import bench  # noqa: E402
from gptest_b200 import _lib  # noqa: E402
from gptest_b200.sweep import natural_params  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
D = int(sys.argv[2]) if len(sys.argv) > 2 else 8
step = int(sys.argv[3]) if len(sys.argv) > 3 else 128
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 8
grid = int(sys.argv[5]) if len(sys.argv) > 5 else 10000
X, y, _, log_hyp = bench.make_c2(N, D)
Z = np.random.default_rng(1).random((grid, D))
kh = natural_params(log_hyp)[0]
h = _lib.default_handle()
n0 = N - step * reps
out = {'N': N, 'D': D, 'step': step, 'appends': reps, 'grid': grid}
for rep in range(2):                                  # first pass warms up allocations
    h.grow_begin(kh, D, capacity=N)
    t0 = time.perf_counter()
    h.grow_append(X[:n0], y[:n0])
    out['initial_fit_ms'] = (time.perf_counter() - t0) * 1e3
    ts = []
    for i in range(reps):
        a = n0 + i * step
        t0 = time.perf_counter()
        v = h.grow_append(X[a:a + step], y[a:a + step])
        ts.append((time.perf_counter() - t0) * 1e3)
        out['last_append_stage_ms'] = h.timings()
    out['append_ms'] = ts
    t0 = time.perf_counter()
    fz, cov = h.grow_predict(Z)
    out['predict_from_factor_ms'] = (time.perf_counter() - t0) * 1e3
h.set_train(X, y)
full = h.gpr_nlml(kh)
t0 = time.perf_counter()
full = h.gpr_nlml(kh)
out['full_refit_ms'] = (time.perf_counter() - t0) * 1e3
pf, pc = h.gpr_predict(kh, Z[:2048])
out['nlml_rel_err'] = abs(v - full) / abs(full)
out['pred_err'] = float(np.abs(fz[:2048] - pf).max())
out['append_flops_model'] = float(N) ** 2 * step
out['speedup_vs_refit'] = out['full_refit_ms'] / float(np.median(ts))
print(json.dumps(out))
