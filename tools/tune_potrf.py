"""Factorisation time vs schedule options: python tools/tune_potrf.py N [N ...]"""
import json
import sys

import torch

sys.path.insert(0, '.')
from gptest_b200 import _lib

sizes = [int(a) for a in sys.argv[1:]] or [16384]
h = _lib.Handle(0)
st = torch.cuda.ExternalStream(h.stream())
out = {}
DEFAULT = dict(nb_tiles=0, lookahead=1, split_tiles=1, small_tile_threshold=296, nb_switch4=64, nb_switch2=24, stagger=1)
VARIANTS = [
    {},
    dict(nb_tiles=4), dict(nb_tiles=2), dict(nb_tiles=1),
    dict(nb_switch4=96, nb_switch2=32), dict(nb_switch4=48, nb_switch2=16), dict(nb_switch4=40, nb_switch2=8),
    dict(small_tile_threshold=600), dict(small_tile_threshold=1200), dict(small_tile_threshold=2400),
    dict(small_tile_threshold=1200, nb_switch4=48, nb_switch2=16),
    dict(stagger=0), dict(split_tiles=0), dict(lookahead=0),
]
for N in sizes:
    M = torch.randn(N, N, dtype=torch.float64, device='cuda')
    K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
    del M
    K2 = torch.empty_like(K)
    for var in VARIANTS:
        opts = dict(DEFAULT)
        opts.update(var)
        for k, v in opts.items():
            h.set_option(k, v)
        best = 1e30
        for it in range(3):
            K2.copy_(K)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            h.potrf_dev(K2.data_ptr(), N, N)
            e1.record(st)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        key = 'N%d %s' % (N, ' '.join('%s=%s' % kv for kv in var.items()) or 'default')
        out[key] = best
        print('%-64s %8.3f ms  %6.2f TFLOP/s' % (key, best, N ** 3 / 3 / best / 1e9), flush=True)
    del K, K2
json.dump(out, open('gpurun_out/tune_potrf.json', 'w'), indent=1)
