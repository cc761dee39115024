"""The DMMA tile kernel in the shape of one trailing update (C -= A B^T, K = nb*128) for ncu:
   python tools/gemm_once.py M N K epi"""
import sys
import torch
sys.path.insert(0, '.')
from gptest_b200 import _lib
M, N, K, epi = (int(a) for a in (sys.argv[1:5] + ['8192', '8192', '512', '1'][len(sys.argv) - 1:]))
h = _lib.Handle(0)
A = torch.randn(M, K, dtype=torch.float64, device='cuda')
B = torch.randn(N, K, dtype=torch.float64, device='cuda')
C = torch.randn(M, N, dtype=torch.float64, device='cuda')
ref = C - A @ B.T if epi else A @ B.T
st = torch.cuda.ExternalStream(h.stream())
torch.cuda.synchronize()
for i in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    h.dgemm_nt_dev(C.data_ptr(), N, A.data_ptr(), K, B.data_ptr(), K, M, N, K, -1.0 if epi else 1.0, 1.0 if epi else 0.0)
    e1.record(st)
    torch.cuda.synchronize()
    if i == 0:
        print('maxerr', (C - ref).abs().max().item())
    ms = e0.elapsed_time(e1)
    print('ms %.4f  TFLOP/s %.2f' % (ms, 2.0 * M * N * K / ms / 1e9))
