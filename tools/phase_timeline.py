"""Where the N=16384 factorisation spends its time, by block-width phase.

The outer block width follows the REMAINING matrix (chol.cu: 8 tiles while >= 96 tile columns remain, 4 down to 64,
2 down to 40, then 1), so the sweep over the last N' columns of a big matrix is the sweep of a fresh N' x N' matrix.
Timing fresh factorisations at the switch points therefore splits the big one into its phases with CUDA events and
no instrumentation: phase(a -> b) = T(a) - T(b), flops (a^3 - b^3)/3.   python tools/phase_timeline.py > profiles/...json"""
import json
import sys
import torch
sys.path.insert(0, '.')
from gptest_b200 import _lib

h = _lib.Handle(0)
st = torch.cuda.ExternalStream(h.stream())
peak = h.microbench(0)
pts = [16384, 12288, 8192, 5120]
T = {}
for N in pts:
    M = torch.randn(N, N, dtype=torch.float64, device='cuda')
    K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
    del M
    K2 = torch.empty_like(K)
    best = 1e30
    for it in range(5):
        K2.copy_(K)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        h.potrf_dev(K2.data_ptr(), N, N)
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    T[N] = best
    del K, K2
out = {"dmma_peak_tflops": peak, "potrf_ms": T, "phases": []}
names = ["8-tile blocks (K=1024)", "4-tile blocks (K=512)", "2-tile blocks (K=256)", "1-tile blocks (K=128), look-ahead tail"]
for i, name in enumerate(names):
    a = pts[i]
    b = pts[i + 1] if i + 1 < len(pts) else 0
    ms = T[a] - (T[b] if b else 0.0)
    fl = (a ** 3 - b ** 3) / 3.0
    out["phases"].append({"phase": name, "columns": "%d -> %d remaining" % (a, b), "ms": ms, "tflops": fl / ms / 1e9,
                          "frac_of_dmma_peak": fl / ms / 1e9 / peak, "ms_at_peak": fl / peak / 1e9})
print(json.dumps(out, indent=1))
