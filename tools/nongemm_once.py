"""One pass over the non-GEMM kernels for ncu: two Newton iterations of C4 (pref_pair / pref_row /
scale_copy_lower / trsv_lt_step / row_dot / pref_finish), a value+gradient fit at N=4096 (grad_trace_kernel), and a
32-problem N=2048 batch with the separate forward substitution (trsv_l_step).  python tools/nongemm_once.py"""
import sys
import numpy as np
sys.path.insert(0, '.')
import bench_configs as cfg
from gptest_b200 import _lib

h = _lib.Handle(0)
X, uvi, y, lh = cfg.make_c4()
D = X.shape[1]
h.set_train(X)
kh = np.r_[np.exp(lh[:D]), np.exp(lh[D]) ** 2]
f, lml, it, tr, jit = h.pref_laplace(uvi, y, kh, sigma=1.0, delta_f=0.0, max_iter=2)
print('c4', it, lml)
X2, y2, Z2, lh2 = cfg.make_c2(n=4096)
h.set_train(X2, y2)
v, g = h.gpr_nlml(cfg.khyp_of(lh2), want_grad=True)
print('grad', v, g[:3])
X5, Y5, lhs = cfg.make_c5()
h.set_train(X5, Y5)
h.set_option('fuse_rhs', 0)
vals, info = h.gpr_nlml_batched(np.array([cfg.khyp_of(l) for l in lhs[:32]]))
print('batched', vals[:2])
