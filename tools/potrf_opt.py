"""Factorisation times under one option of the library: python tools/potrf_opt.py <option> <value> [<value> ...]"""
import sys
import torch
sys.path.insert(0, '.')
from gptest_b200 import _lib
opt, vals = sys.argv[1], [int(v) for v in sys.argv[2:]]
h = _lib.Handle(0)
st = torch.cuda.ExternalStream(h.stream())
for N in (2048, 4096, 8192, 16384):
    M = torch.randn(N, N, dtype=torch.float64, device='cuda')
    K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
    del M
    K2 = torch.empty_like(K)
    res = {}
    for rep in range(2):
        for v in vals:
            h.set_option(opt, v)
            ts = []
            for it in range(5):
                K2.copy_(K)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                h.potrf_dev(K2.data_ptr(), N, N)
                e1.record(st)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res.setdefault(v, []).append(round(min(ts), 4))
    print(N, opt, res, flush=True)
    del K, K2
