"""Round-2 retune of the schedule knobs with the faster panel chain: python tools/tune2.py [N ...]"""
import sys
import itertools
import torch
sys.path.insert(0, '.')
from gptest_b200 import _lib

sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384]
h = _lib.Handle(0)
st = torch.cuda.ExternalStream(h.stream())
defaults = dict(small_tile_threshold=2400, nb_switch2=24, nb_switch4=64, pdl_tail=0, thin_tile_max=74, pdl_max_tiles=40)
trials = [dict()] + [dict(nb_switch2=v) for v in (12, 16, 20, 32)] + [dict(pdl_max_tiles=24), dict(thin_tile_max=37), dict(pdl_max_tiles=24, thin_tile_max=37),
          dict(nb_switch2=16, pdl_max_tiles=24), dict(small_tile_threshold=1800), dict(pdl_tail=1), dict(tri_skip=0)]
for N in sizes:
    M = torch.randn(N, N, dtype=torch.float64, device='cuda')
    K = M @ M.T / N + torch.eye(N, dtype=torch.float64, device='cuda')
    del M
    K2 = torch.empty_like(K)
    for tr in trials:
        for k, v in dict(defaults, tri_skip=1).items():
            h.set_option(k, tr.get(k, v))
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
        print('N', N, tr or 'defaults', round(best, 3), 'ms', round(N ** 3 / 3 / best / 1e9, 2), 'TFLOP/s', flush=True)
    del K, K2
