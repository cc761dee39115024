"""Chunked multi-stream trailing update on/off: python tools/dag_ab.py [N ...]"""
import json
import sys

import torch

sys.path.insert(0, '.')
from gptest_b200 import _lib

sizes = [int(a) for a in sys.argv[1:]] or [16384, 8192, 4096]
h = _lib.Handle(0)
st = torch.cuda.ExternalStream(h.stream())
VARIANTS = [dict(), dict(dag_min_width=2), dict(dag_min_width=1), dict(dag_min_width=2, nb_switch2=16), dict(dag_min_width=2, nb_switch4=48),
            dict(dag_min_width=2, dag_big_tiles=0), dict(dag_min_width=1, nb_switch2=32), dict()]
DEFAULT = dict(dag_streams=4, stagger=1, dag_big_tiles=1, nb_switch8=96, dag_min_width=4, nb_switch4=64, nb_switch2=40, dag_min_tiles=8)
out = {}
for N in sizes:
    M = torch.randn(N, N, dtype=torch.float64, device='cuda')
    K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
    del M
    K2 = torch.empty_like(K)
    ref = None
    for var in VARIANTS:
        opts = dict(DEFAULT)
        opts.update(var)
        for k, v in opts.items():
            h.set_option(k, v)
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
        key = 'N%d %s' % (N, ' '.join('%s=%s' % kv for kv in var.items()))
        out[key] = (best, bool(torch.equal(L, ref)))
        print('%-60s %8.3f ms %6.2f TF  bitwise=%s' % (key, best, N ** 3 / 3 / best / 1e9, out[key][1]), flush=True)
    del K, K2, ref, L
json.dump(out, open('gpurun_out/dag_ab.json', 'w'), indent=1)
