"""Programmatic dependent launch on/off: python tools/pdl_ab.py
potrf at several N (CUDA events), one preference Laplace call (n=4096, 25 iterations) and a 256 x 2048 sweep."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
import bench
from gptest_b200 import _lib

h = _lib.Handle(0)
st = torch.cuda.ExternalStream(h.stream())
out = {}
for N in (1024, 2048, 4096, 8192, 16384):
    M = torch.randn(N, N, dtype=torch.float64, device='cuda')
    K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
    del M
    K2 = torch.empty_like(K)
    ref = None
    for mode in (0, 1, 2, 0, 1):
        h.set_option('pdl', mode)
        best = 1e30
        for it in range(4):
            K2.copy_(K)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            h.potrf_dev(K2.data_ptr(), N, N)
            e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        L = torch.tril(K2)
        if ref is None:
            ref = L.clone()
        same = bool(torch.equal(L, ref))
        out.setdefault('potrf_N%d' % N, []).append((mode, round(best, 4), same))
    print('potrf N', N, out['potrf_N%d' % N], flush=True)
    del K, K2, ref, L

rng = np.random.default_rng(0)
n, D, P = 4096, 6, 32768
Xp = rng.random((n, D))
uvi = rng.integers(0, n, (P, 2))
bad = uvi[:, 0] == uvi[:, 1]
uvi[bad, 1] = (uvi[bad, 0] + 1) % n
w = rng.standard_normal(D)
lat = np.sin(2 * np.pi * Xp @ w / np.abs(w).sum() + np.pi / 4) + 0.2
yp = np.where(lat[uvi[:, 1]] + 0.05 * rng.standard_normal(P) > lat[uvi[:, 0]] + 0.05 * rng.standard_normal(P), 1.0, -1.0)
h.set_train(Xp)
for mode in (0, 1, 2, 0, 1):
    h.set_option('pdl', mode)
    ts = []
    for rep in range(2):
        t0 = time.perf_counter()
        f, lml, iters, trace, jit = h.pref_laplace(uvi, yp, np.r_[[0.5] * D, 1.0], sigma=1.0, delta_f=1e-6, max_iter=25)
        ts.append((time.perf_counter() - t0) * 1e3)
    out.setdefault('pref_25iter_ms', []).append((mode, round(min(ts), 3), float(lml)))
print('pref', out['pref_25iter_ms'], flush=True)

X, Y, lhs = bench.make_c5(2048, 1024)
kh = np.array([bench.khyp_of(l) for l in lhs[:256]])
h.set_train(X, Y)
for mode in (0, 1, 2, 0, 1):
    h.set_option('pdl', mode)
    h.gpr_nlml_batched(kh)
    t0 = time.perf_counter()
    vals, info = h.gpr_nlml_batched(kh)
    out.setdefault('sweep256x2048_ms', []).append((mode, round((time.perf_counter() - t0) * 1e3, 3), float(vals.sum())))
print('sweep', out['sweep256x2048_ms'], flush=True)
json.dump(out, open('gpurun_out/pdl_ab.json', 'w'), indent=1)
