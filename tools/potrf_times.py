"""Factorisation times of the library selected by $GPB200_LIB (default: the in-tree build): python tools/potrf_times.py [N ...]"""
import sys
import torch
sys.path.insert(0, '.')
from gptest_b200 import _lib
sizes = [int(a) for a in sys.argv[1:]] or [1024, 4096, 8192, 16384]
h = _lib.Handle(0)
st = torch.cuda.ExternalStream(h.stream())
out = {}
for N in sizes:
    M = torch.randn(N, N, dtype=torch.float64, device='cuda')
    K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
    del M
    K2 = torch.empty_like(K)
    ts = []
    for it in range(6):
        K2.copy_(K)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        h.potrf_dev(K2.data_ptr(), N, N)
        e1.record(st)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out[N] = round(min(ts), 4)
    del K, K2
print(_lib.LIB_PATH.split('/')[-1], out, flush=True)
