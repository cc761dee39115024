"""Factorisation time vs schedule options: python tools/tune_potrf.py N [N ...]"""
import itertools
import json
import sys

import torch

sys.path.insert(0, '.')
from gptest_b200 import _lib

sizes = [int(a) for a in sys.argv[1:]] or [16384]
h = _lib.Handle(0)
st = torch.cuda.ExternalStream(h.stream())
out = {}
for N in sizes:
    M = torch.randn(N, N, dtype=torch.float64, device='cuda')
    K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
    del M
    K2 = torch.empty_like(K)
    grid = list(itertools.product((0, 1, 2, 4, 8), (1,), (1,), (296,)))
    grid += [(4, 1, 0, 296), (4, 1, 1, 600), (8, 1, 1, 600), (2, 1, 1, 600), (4, 1, 1, 148), (4, 0, 1, 296)]
    for nb, la, split, thr in grid:
        h.set_option('nb_tiles', nb)
        h.set_option('lookahead', la)
        h.set_option('split_tiles', split)
        h.set_option('small_tile_threshold', thr)
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
        key = 'N%d nb%d la%d split%d thr%d' % (N, nb, la, split, thr)
        out[key] = best
        print('%-34s %8.3f ms  %6.2f TFLOP/s' % (key, best, N ** 3 / 3 / best / 1e9), flush=True)
    del K, K2
json.dump(out, open('gpurun_out/tune_potrf.json', 'w'), indent=1)
