"""First-light measurements on a B200: FP64 pipe rates, cuBLAS/cuSOLVER bars, stage timings."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from gptest_b200 import _lib

out = {}
h = _lib.Handle(0)
out['dmma_tflops'] = h.microbench(0)
out['dfma_tflops'] = h.microbench(1)
print(out, flush=True)


def cuda_time(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


n = 8192
A = torch.randn(n, n, dtype=torch.float64, device='cuda')
B = torch.randn(n, n, dtype=torch.float64, device='cuda')
ms = cuda_time(lambda: A @ B.T)
out['cublas_dgemm_8192_tflops'] = 2 * n ** 3 / ms / 1e9
print(out, flush=True)
C = torch.empty(n, n, dtype=torch.float64, device='cuda')
st = torch.cuda.ExternalStream(h.stream())


def mine():
    h.dgemm_nt_dev(C.data_ptr(), n, A.data_ptr(), n, B.data_ptr(), n, n, n, n, 1.0, 0.0)


def timed_on_handle_stream(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        fn()
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


ms = timed_on_handle_stream(mine)
out['dmma_gemm_8192_tflops'] = 2 * n ** 3 / ms / 1e9
out['dmma_gemm_8192_maxerr'] = (C - A @ B.T).abs().max().item()
print(out, flush=True)
del A, B, C

for N in (4096, 16384):
    M = torch.randn(N, N, dtype=torch.float64, device='cuda')
    K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
    del M
    K2 = K.clone()
    ms = cuda_time(lambda: torch.linalg.cholesky(K), iters=2)
    out['cusolver_potrf_%d_ms' % N] = ms
    out['cusolver_potrf_%d_tflops' % N] = N ** 3 / 3 / ms / 1e9
    for la, nb in ((0, 2), (1, 2), (1, 1), (1, 4)):
        h.set_option('lookahead', la)
        h.set_option('nb_tiles', nb)
        best = 1e30
        for it in range(2):
            K2.copy_(K)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            info = h.potrf_dev(K2.data_ptr(), N, N)
            e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out['gpb_potrf_%d_la%d_nb%d_ms' % (N, la, nb)] = best
        out['gpb_potrf_%d_la%d_nb%d_tflops' % (N, la, nb)] = N ** 3 / 3 / best / 1e9
    Lref = torch.linalg.cholesky(K)
    out['gpb_potrf_%d_maxerr_vs_cusolver' % N] = (torch.tril(K2) - Lref).abs().max().item()
    print(out, flush=True)
    del K, K2, Lref
h.set_option('lookahead', 1)
h.set_option('nb_tiles', 2)

rng = np.random.default_rng(0)
n, d = 16384, 8
X = rng.random((n, d))
y = np.sin(X @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(n)
kh = np.r_[[0.5] * d, 1.0, 0.01]
h.set_train(X, y)
for _ in range(3):
    t0 = time.perf_counter()
    v = h.gpr_nlml(kh)
    t1 = time.perf_counter()
    out['nlml_16384'] = v
    out['nlml_16384_wall_ms'] = (t1 - t0) * 1e3
    out['nlml_16384_stages'] = h.timings()
print(out, flush=True)
Kd = torch.empty(n, n, dtype=torch.float64, device='cuda')
for _ in range(3):
    h.kxx_dev(kh, Kd.data_ptr())
    out['kxx_full_16384_ms'] = h.timings()['kbuild_ms']
out['kxx_full_16384_GBs'] = (8.0 * n * n + 8 * n * d) / out['kxx_full_16384_ms'] / 1e6
Z = rng.random((1024, d))
for _ in range(2):
    fz, cov = h.gpr_predict(kh, Z)
    out['predict_16384_stages'] = h.timings()
json.dump(out, open('gpurun_out/first_light.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
