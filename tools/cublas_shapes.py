"""cuBLAS DGEMM at the shapes of the trailing update (vendor bar for the DMMA tile kernel)."""
import torch
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
M = N = 8192
for K in (128, 256, 512, 1024, 8192):
    A = torch.randn(M, K, dtype=torch.float64, device='cuda'); B = torch.randn(N, K, dtype=torch.float64, device='cuda')
    C = torch.randn(M, N, dtype=torch.float64, device='cuda')
    ms = t(lambda: torch.addmm(C, A, B.T, alpha=-1.0, beta=1.0, out=C))
    print('cuBLAS addmm M=N=8192 K=%5d: %.3f ms  %.2f TFLOP/s' % (K, ms, 2.0 * M * N * K / ms / 1e9))
