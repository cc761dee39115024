"""A/B of a handle option on the factorisation: python tools/variant_ab.py OPTION V0 V1 [N ...]"""
import sys
import numpy as np
import torch

sys.path.insert(0, '.')
from gptest_b200 import _lib

opt, v0, v1 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
sizes = [int(a) for a in sys.argv[4:]] or [128, 1024, 2048, 4096, 8192, 16384]
h = _lib.Handle(0)
st = torch.cuda.ExternalStream(h.stream())
for N in sizes:
    M = torch.randn(N, N, dtype=torch.float64, device='cuda')
    K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
    del M
    K2 = torch.empty_like(K)
    res = {}
    outs = {}
    for mode in (v0, v1, v0, v1):
        h.set_option(opt, mode)
        best = 1e30
        for it in range(4):
            K2.copy_(K)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            info = h.potrf_dev(K2.data_ptr(), N, N)
            e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res.setdefault(mode, []).append(round(best, 4))
        outs[mode] = torch.tril(K2).clone()
    err = float((outs[v1] - torch.linalg.cholesky(K)).abs().max())
    print('N', N, res, 'bitwise equal', bool(torch.equal(outs[v0], outs[v1])), 'max diff', float((outs[v0] - outs[v1]).abs().max()),
          'vs torch', err, 'info', info, flush=True)
    del K, K2, outs
rng = np.random.default_rng(0)
X = rng.random((3000, 4)); y = np.sin(X.sum(1)); Z = rng.random((200, 4))
kh = np.array([0.5] * 4 + [1.0, 0.01])
h.set_train(X, y)
out = {}
for mode in (v0, v1):
    h.set_option(opt, mode)
    out[mode] = (h.gpr_nlml(kh), h.gpr_predict(kh, Z), h.gpr_nlml(kh, want_grad=True)[1])
print('nlml', out[v0][0], out[v1][0], 'pred diff', float(np.abs(out[v0][1][0] - out[v1][1][0]).max()),
      float(np.abs(out[v0][1][1] - out[v1][1][1]).max()), 'grad diff', float(np.abs(out[v0][2] - out[v1][2]).max()))
A = np.eye(300); A[150, 150] = -1.0
for mode in (v0, v1):
    h.set_option(opt, mode)
    try:
        h.potrf(A); print('no error?!')
    except np.linalg.LinAlgError as e:
        print(mode, 'LinAlgError', e)
