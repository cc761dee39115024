"""The DMMA tile kernel in the shape of one trailing update (C -= A B^T, K = nb*128) for ncu / timing:
   python tools/gemm_once.py M N K epi [split_tiles]
   epi: 0 -> C = A B^T, 1 -> C -= A B^T, 2 -> k loop only, nothing stored (measurement)"""
import sys
import torch
sys.path.insert(0, '.')
from gptest_b200 import _lib
M, N, K, epi = (int(a) for a in (sys.argv[1:5] + ['8192', '8192', '512', '1'][len(sys.argv) - 1:]))
h = _lib.Handle(0)
if len(sys.argv) > 5:
    h.set_option('split_tiles', int(sys.argv[5]))
if len(sys.argv) > 6:
    h.set_option('stagger', int(sys.argv[6]))
A = torch.randn(M, K, dtype=torch.float64, device='cuda')
B = torch.randn(N, K, dtype=torch.float64, device='cuda')
C = torch.randn(M, N, dtype=torch.float64, device='cuda')
ref = C - A @ B.T if epi == 1 else A @ B.T
alpha, beta = {0: (1.0, 0.0), 1: (-1.0, 1.0), 2: (0.0, 0.0)}[epi]
st = torch.cuda.ExternalStream(h.stream())
torch.cuda.synchronize()
for i in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    h.dgemm_nt_dev(C.data_ptr(), N, A.data_ptr(), K, B.data_ptr(), K, M, N, K, alpha, beta)
    e1.record(st)
    torch.cuda.synchronize()
    if i == 0 and epi != 2:
        print('maxerr', (C - ref).abs().max().item())
    ms = e0.elapsed_time(e1)
    print('ms %.4f  TFLOP/s %.2f' % (ms, 2.0 * M * N * K / ms / 1e9))
