"""Throughput of a SMALL batch of BIG fits (independent hyper-parameter vectors on one training set):
    python tools/batch_big.py N B [la_max_batch ...]
Compares B sequential gpr_nlml calls with one gpr_nlml_batched call of B problems."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import bench  # noqa: E402
from gptest_b200 import _lib  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
las = [int(a) for a in sys.argv[3:]] or [1, 8]
from gptest_b200.sweep import natural_params  # noqa: E402
X, y, _, log_hyp = bench.make_c2(N, 8)
kh = natural_params(log_hyp)[0]
h = _lib.default_handle()
h.set_train(X, y)
khs = np.array([kh * (1.0 + 0.01 * i) for i in range(B)])
khs[:, -1] = kh[-1]
out = {'N': N, 'B': B}
ref = np.array([h.gpr_nlml(k) for k in khs])
t0 = time.perf_counter()
for _ in range(2):
    for k in khs:
        h.gpr_nlml(k)
out['sequential_ms_per_fit'] = (time.perf_counter() - t0) / (2 * B) * 1e3
for la in las:
    h.set_option('la_max_batch', la)
    v, info = h.gpr_nlml_batched(khs)
    assert not info.any()
    err = float(np.max(np.abs(v - ref) / np.abs(ref)))
    t0 = time.perf_counter()
    for _ in range(2):
        h.gpr_nlml_batched(khs)
    out['batched_la%d_ms_per_fit' % la] = (time.perf_counter() - t0) / (2 * B) * 1e3
    out['batched_la%d_rel_err' % la] = err
print(json.dumps(out))
