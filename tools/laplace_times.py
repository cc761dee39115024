"""C3 / C4 at full size, wall time per call and per iteration (GPB_DEBUG_LOOP=1 prints the graph build cost): python tools/laplace_times.py"""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import bench_configs as cfg
from gptest_b200 import _lib
h = _lib.Handle(0)
Xc, yc, _, lhc = cfg.make_c3()
D = Xc.shape[1]
h.set_train(Xc)
for rep in range(3):
    t0 = time.perf_counter()
    f, lml, it, tr, jit = h.gpc_laplace(yc, np.r_[np.exp(lhc[:D]), np.exp(lhc[D]) ** 2], link=0, delta_f=1e-6)
    print('c3 %.2f ms, %d iterations' % ((time.perf_counter() - t0) * 1e3, it), h.timings(), flush=True)
Xp, uvi, yp, lhp = cfg.make_c4()
D = Xp.shape[1]
h.set_train(Xp)
for cap in (500, 20, 20):
    t0 = time.perf_counter()
    f, lml, it, tr, jit = h.pref_laplace(uvi, yp, np.r_[np.exp(lhp[:D]), np.exp(lhp[D]) ** 2], sigma=1.0, delta_f=1e-6, max_iter=cap)
    dt = (time.perf_counter() - t0) * 1e3
    tm = h.timings()
    print('c4 %.2f ms, %d iterations, setup %.2f ms, loop %.2f ms = %.3f ms/iteration' % (dt, it, tm['kbuild_ms'], tm['factor_ms'], tm['factor_ms'] / it), flush=True)
