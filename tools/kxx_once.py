"""Full symmetric covariance assembly at N (D=8) for ncu: python tools/kxx_once.py N"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gptest_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = 8
rng = np.random.default_rng(0)
h = _lib.Handle(0)
h.set_train(rng.random((n, d)))
K = torch.empty((n, n), dtype=torch.float64, device='cuda')
kh = np.r_[[0.5] * d, 1.0, 0.01]
for i in range(4):
    h.kxx_dev(kh, K.data_ptr())
    ms = h.timings()['kbuild_ms']
    print('ms %.4f GB/s %.1f' % (ms, (8.0 * n * n + 8 * n * d) / ms / 1e6))
