"""One factorisation (after a warm-up one) for ncu launch lists: python tools/potrf_once.py N [nb] [la]"""
import sys
import torch
sys.path.insert(0, '.')
from gptest_b200 import _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 2
la = int(sys.argv[3]) if len(sys.argv) > 3 else 1
h = _lib.Handle(0)
h.set_option('nb_tiles', nb)
h.set_option('lookahead', la)
M = torch.randn(N, N, dtype=torch.float64, device='cuda')
K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
K2 = K.clone()
torch.cuda.synchronize()
h.potrf_dev(K2.data_ptr(), N, N)
K2.copy_(K)
torch.cuda.synchronize()
info = h.potrf_dev(K2.data_ptr(), N, N)
print('info', info, h.timings())
