"""GPU parity tests of the regression path: CUDA (through the C ABI) vs the CPU oracle.

Tolerances (BASELINE.json north_star): relative 1e-9 on posterior mean and variance, 1e-8 on
the log marginal likelihood.  Where the reference's own inv()-based arithmetic is noisier than
that the comparison is made against the Cholesky form of the oracle and the
reference-vs-oracle gap is asserted separately, so nothing is hidden.
"""
import json
import os

import numpy as np
import pytest

from oracle import gpr_oracle

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'gpr_kat.json')))


def khyp_of(log_hyp):
    ell, sf2, sn2 = gpr_oracle.split_hyp(log_hyp)
    return np.concatenate([ell, [sf2, sn2]])


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


# ---------------------------------------------------------------------------------------
# the DMMA tile kernel and the factorisation on their own
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(128, 128, 128), (256, 384, 512), (1024, 512, 256)])
def test_dgemm_nt_matches_fp64_matmul(handle, M, N, K):
    import torch
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    A = torch.randn(M, K, dtype=torch.float64, device='cuda', generator=g)
    B = torch.randn(N, K, dtype=torch.float64, device='cuda', generator=g)
    C = torch.full((M, N), float('nan'), dtype=torch.float64, device='cuda')
    torch.cuda.synchronize()
    handle.dgemm_nt_dev(C.data_ptr(), N, A.data_ptr(), K, B.data_ptr(), K, M, N, K, 1.0, 0.0)
    torch.cuda.synchronize()
    ref = A @ B.T
    err = (C - ref).abs().max().item()
    assert err <= 1e-12 * K, err
    C0 = torch.randn(M, N, dtype=torch.float64, device='cuda', generator=g)
    C1 = C0.clone()
    torch.cuda.synchronize()
    handle.dgemm_nt_dev(C1.data_ptr(), N, A.data_ptr(), K, B.data_ptr(), K, M, N, K, -1.0, 1.0)
    torch.cuda.synchronize()
    err = (C1 - (C0 - ref)).abs().max().item()
    assert err <= 1e-12 * K, err


@pytest.mark.parametrize("n", [5, 128, 200, 384, 1000, 2048])
def test_potrf_matches_numpy(handle, n):
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    A = M @ M.T / n + np.eye(n)
    L = handle.potrf(A)
    Lref = np.linalg.cholesky(A)
    assert np.array_equal(np.triu(L, 1), np.zeros_like(L))
    assert np.abs(L - Lref).max() < 1e-12
    resid = np.abs(L @ L.T - A).max() / np.abs(A).max()
    assert resid < 1e-13, resid


def test_potrf_lookahead_and_block_width_are_bitwise_equivalent(handle):
    """the schedule must not change the arithmetic: every tile sees the same operations"""
    rng = np.random.default_rng(7)
    n = 1536
    M = rng.standard_normal((n, n))
    A = M @ M.T / n + np.eye(n)
    outs = []
    try:
        for la, nb in [(0, 1), (1, 1), (0, 2), (1, 2), (1, 4)]:
            handle.set_option('lookahead', la)
            handle.set_option('nb_tiles', nb)
            outs.append(handle.potrf(A))
    finally:
        handle.set_option('lookahead', 1)
        handle.set_option('nb_tiles', 0)
    # same block width => identical bits with and without look-ahead
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(outs[2], outs[3])
    for o in outs[1:]:
        assert np.abs(o - outs[0]).max() < 1e-12


def test_chunked_schedule_and_dependent_launch_are_bitwise_equivalent(handle):
    """The chunked multi-stream trailing update (chol.cu) and programmatic dependent launch only reorder launches:
    same tiles, same per-tile order of operations.  Forced on at a small size (dag_min_tiles, block widths 8/4/2/1
    all exercised through the switch points) and compared with the one-launch schedule."""
    rng = np.random.default_rng(11)
    n = 2560                                                   # 20 tile columns
    M = rng.standard_normal((n, n))
    A = M @ M.T / n + np.eye(n)
    opts = dict(dag_min_tiles=8, nb_switch8=16, nb_switch4=10, nb_switch2=4, small_tile_threshold=40)
    outs = {}
    try:
        for k, v in opts.items():
            handle.set_option(k, v)
        for name, extra in [('one_launch', dict(dag_streams=0)), ('chunked', dict(dag_streams=3)),
                            ('chunked_narrow', dict(dag_streams=3, dag_min_width=2)),
                            ('chunked_small_tiles', dict(dag_streams=5, dag_big_tiles=0)),
                            ('no_pdl', dict(dag_streams=3, pdl=0)), ('pdl_all', dict(dag_streams=3, pdl=2))]:
            handle.set_option('dag_streams', 4); handle.set_option('dag_min_width', 4)
            handle.set_option('dag_big_tiles', 1); handle.set_option('pdl', 1)
            for k, v in extra.items():
                handle.set_option(k, v)
            outs[name] = handle.potrf(A)
    finally:
        for k, v in dict(dag_min_tiles=72, nb_switch8=96, nb_switch4=64, nb_switch2=40, small_tile_threshold=2400,
                         dag_streams=4, dag_min_width=4, dag_big_tiles=1, pdl=1).items():
            handle.set_option(k, v)
    ref = np.linalg.cholesky(A)
    assert np.abs(outs['one_launch'] - ref).max() < 1e-11
    for name in ('chunked', 'chunked_narrow', 'no_pdl', 'pdl_all'):
        assert np.array_equal(outs[name], outs['one_launch']), name
    # a different tile shape changes the summation order inside a tile: equal to rounding only
    assert np.abs(outs['chunked_small_tiles'] - outs['one_launch']).max() < 1e-12


def test_full_size_schedule_is_race_free(handle):
    """The default schedule at a size where all of it is active (chunked multi-stream updates with 8- and 4-tile
    blocks, look-ahead tail, chain on the panel stream): repeated runs must give the same bits as each other and as
    the plain one-launch-per-step order (no look-ahead, no chunks) with the same block widths."""
    import torch
    n = 14336                                                  # 112 tile columns: 8-tile, 4-tile, 2-tile and 1-tile phases
    g = torch.Generator(device='cuda').manual_seed(5)
    M = torch.randn(n, 512, dtype=torch.float64, device='cuda', generator=g)
    K = M @ M.T / 512 + torch.eye(n, dtype=torch.float64, device='cuda')
    del M
    W = torch.empty_like(K)
    outs = []
    try:
        for rep in range(3):
            W.copy_(K)
            torch.cuda.synchronize()
            assert handle.potrf_dev(W.data_ptr(), n, n) == 0
            outs.append(torch.tril(W).clone())
        handle.set_option('lookahead', 0)
        W.copy_(K)
        torch.cuda.synchronize()
        assert handle.potrf_dev(W.data_ptr(), n, n) == 0
        plain = torch.tril(W).clone()
    finally:
        handle.set_option('lookahead', 1)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert torch.equal(outs[0], plain)
    # and it is a factor of K (probe columns)
    cols = torch.tensor([0, 1, 127, 128, 4097, n - 129, n - 1], device='cuda')
    R = outs[0] @ outs[0][cols].T - K[:, cols]
    assert float(R.abs().max()) < 1e-11 * float(K.abs().max())


def test_potrf_reports_not_positive_definite(handle):
    A = np.eye(300)
    A[150, 150] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        handle.potrf(A)


# ---------------------------------------------------------------------------------------
# covariance assembly (GPr.py:99-110)
# ---------------------------------------------------------------------------------------
def test_kxx_kat1_golden(handle):
    k = GOLD['kat1']
    handle.set_train(np.array(k['x']))
    K = handle.kxx(khyp_of(k['log_hyp']))
    Kref = np.array(k['K'])
    assert K.shape == Kref.shape
    assert np.abs(K - Kref).max() <= 4e-16 * 1.01
    assert np.array_equal(K, K.T)
    assert K[0, 0] == Kref[0, 0]


@pytest.mark.parametrize("n,d", [(1, 1), (63, 2), (64, 1), (300, 3), (1024, 8), (1500, 20)])
def test_kxx_matches_oracle(handle, n, d):
    rng = np.random.default_rng(n * 31 + d)
    x = rng.random((n, d))
    lh = np.log(np.r_[0.3 + rng.random(d), 1.3, 0.1])
    handle.set_train(x)
    K = handle.kxx(khyp_of(lh))
    Kref = gpr_oracle.kxx(lh, x)
    sf2 = np.exp(lh[d]) ** 2
    assert np.abs(K - Kref).max() <= 2e-14 * sf2
    assert np.array_equal(K, K.T)


@pytest.mark.parametrize("n,m,d", [(20, 100, 1), (300, 77, 3), (1024, 256, 8)])
def test_kxz_matches_oracle(handle, n, m, d):
    rng = np.random.default_rng(n + m + d)
    x, z = rng.random((n, d)), rng.random((m, d))
    lh = np.log(np.r_[0.3 + rng.random(d), 0.9, 0.1])
    handle.set_train(x)
    Kxz = handle.kxz(khyp_of(lh), z)
    ref = gpr_oracle.kxz(lh, x, z)
    assert Kxz.shape == (n, m)
    assert np.abs(Kxz - ref).max() <= 2e-14


def test_sqdist_matches_reference_form(handle):
    rng = np.random.default_rng(5)
    a, b = rng.random((130, 3)), rng.random((70, 3))
    handle.set_train(a)
    got = handle.sqdist(b)
    assert np.abs(got - gpr_oracle.sqdist_expanded(a, b)).max() < 1e-14


# ---------------------------------------------------------------------------------------
# likelihood and prediction (GPr.py:45-69)
# ---------------------------------------------------------------------------------------
def test_nlml_kat1_kat2_golden(handle):
    k = GOLD['kat1']
    handle.set_train(np.array(k['x']), np.array(k['y']))
    v = handle.gpr_nlml(khyp_of(k['log_hyp']))
    assert abs(v - k['nlml']) <= 1e-8 * abs(k['nlml'])
    rng = np.random.default_rng(0)
    x = rng.random(1024)
    y = np.sin(6 * x) + 0.1 * rng.standard_normal(1024)
    k2 = GOLD['kat2']
    handle.set_train(x, y)
    v = handle.gpr_nlml(khyp_of(k2['log_hyp']))
    assert abs(v - k2['nlml']) <= 1e-8 * abs(k2['nlml'])
    k2b = GOLD['kat2b']
    handle.set_train(x[:512], y[:512])
    v = handle.gpr_nlml(khyp_of(k2b['log_hyp']))
    assert abs(v - k2b['nlml']) <= 1e-8 * abs(k2b['nlml'])


def test_predict_kat1_golden(handle):
    k = GOLD['kat1']
    handle.set_train(np.array(k['x']), np.array(k['y']))
    fz, cov = handle.gpr_predict(khyp_of(k['log_hyp']), np.array(k['z']))
    assert rel(fz, k['mean']) < 1e-9
    # variance: 1e-9 relative to the prior variance sf2 = 1 (the reference's own inv() noise on
    # these near-zero variances is larger than 1e-9 of their value)
    assert np.abs(cov - np.array(k['var'])).max() < 1e-9


@pytest.mark.parametrize("n,m,d", [(777, 50, 8), (2048, 300, 8), (1000, 129, 2)])
def test_nlml_and_predict_match_oracle(handle, n, m, d):
    rng = np.random.default_rng(n)
    x = rng.random((n, d))
    w = rng.standard_normal(d)
    y = np.sin(x @ w) + 0.1 * rng.standard_normal(n)
    z = rng.random((m, d))
    lh = np.log([0.5] * d + [1.0, 0.1])
    handle.set_train(x, y)
    v = handle.gpr_nlml(khyp_of(lh))
    ref = float(gpr_oracle.nlml(lh, x, y)[0, 0])
    assert abs(v - ref) <= 1e-8 * abs(ref), (v, ref)
    fz, cov = handle.gpr_predict(khyp_of(lh), z)
    rm, rv = gpr_oracle.predict(lh, x, y, z)
    cm, cv = gpr_oracle.predict_chol(lh, x, y, z)
    # the oracle's two formulations bound what "the reference's answer" means at this conditioning
    ref_gap_m = rel(rm, cm)
    assert rel(fz, cm) < 1e-9, rel(fz, cm)
    assert rel(fz, rm) < max(1e-9, 3 * ref_gap_m)
    assert np.abs(cov - cv).max() < 1e-9
    assert np.abs(cov - rv).max() < max(1e-9, 3 * np.abs(rv - cv).max())


def test_batched_nlml_equals_single_calls(handle):
    rng = np.random.default_rng(11)
    n, d, B = 700, 2, 9
    x = 100 * rng.random((n, d))
    y = np.sin(x[:, 0] / 20) + 0.25 * rng.standard_normal(n)
    handle.set_train(x, y)
    lhs = np.log(np.c_[20 * np.exp(0.3 * rng.standard_normal((B, d))), np.sqrt(10) * np.ones(B), np.ones(B)])
    kh = np.array([khyp_of(l) for l in lhs])
    vals, info = handle.gpr_nlml_batched(kh)
    assert (info == 0).all()
    singles = np.array([handle.gpr_nlml(k) for k in kh])
    # the batched schedule may pick other tile / block sizes than a single call: same math, other rounding
    assert rel(vals, singles) < 1e-13
    refs = np.array([float(gpr_oracle.nlml(l, x, y)[0, 0]) for l in lhs])
    assert rel(vals, refs) < 1e-8
    try:
        handle.set_option('batch_chunk', 4)       # ragged chunks: 4 + 4 + 1
        vals2, _ = handle.gpr_nlml_batched(kh)
    finally:
        handle.set_option('batch_chunk', 0)
    # (a chunk of one problem takes the single-matrix schedule, i.e. other block widths: equal to rounding)
    assert rel(vals, vals2) < 1e-13


# ---------------------------------------------------------------------------------------
# the drop-in module surface (GPr.py:16-110)
# ---------------------------------------------------------------------------------------
def test_dropin_module_reproduces_demo(handle):
    from gptest_b200 import GPr
    k = GOLD['kat1']
    x, y, z = np.array(k['x']), np.array(k['y']), np.array(k['z'])
    lh = np.array(k['log_hyp'])
    gp = GPr.GaussianProcess(lh, 0, 0, "SE", "zero", "zero", x, y)
    out = gp.compute_likelihood(lh)
    assert out.shape == (1, 1)
    assert abs(out[0, 0] - k['nlml']) <= 1e-8 * abs(k['nlml'])
    fz, cov = gp.compute_prediction(z)
    assert fz.shape == (100,) and cov.shape == (100,)
    assert rel(fz, k['mean']) < 1e-9
    assert gp.covFun.sf2 == np.exp(lh)[1] ** 2 and gp.meanFun.y == 0
    assert GPr.GaussianProcess(lh, 0, 0, "nope", "nope", "nope", x, y).covFun == []
    # the optimiser loop of GP_regression_demo.py:44 runs unmodified
    import scipy.optimize as op
    opt = op.fmin(gp.compute_likelihood, lh, disp=False)
    kb = GOLD['kat1b']
    assert abs(float(gp.compute_likelihood(opt)[0, 0]) - kb['nlml']) < 1e-5
    gp2 = GPr.GaussianProcess(opt, 0, 0, "SE", "zero", "zero", x, y)
    fz2, _ = gp2.compute_prediction(z)
    assert np.abs(fz2 - np.array(kb['mean'])).max() < 1e-3


# ---------------------------------------------------------------------------------------
# full size (BASELINE config 2: N = 16384, D = 8) through size-independent properties
# ---------------------------------------------------------------------------------------
def test_full_size_fit_properties(handle):
    import torch
    rng = np.random.default_rng(0)
    n, d, m = 16384, 8, 1024
    X = rng.random((n, d))
    w = rng.standard_normal(d)
    y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
    Z = rng.random((m, d))
    lh = np.log([0.5] * d + [1.0, 0.1])
    kh = khyp_of(lh)
    handle.set_train(X, y)
    v = handle.gpr_nlml(kh)
    # independent check: assemble K with our kernel, factor it with cuSOLVER through torch
    K = torch.empty((n, n), dtype=torch.float64, device='cuda')
    handle.kxx_dev(kh, K.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(K, K.T)
    sub = K[:256, :256].cpu().numpy()
    assert np.abs(sub - gpr_oracle.kxx(lh, X[:256])).max() < 2e-14
    Lt = torch.linalg.cholesky(K)
    yt = torch.from_numpy(y).cuda()
    zt = torch.linalg.solve_triangular(Lt, yt[:, None], upper=False)[:, 0]
    ref = (0.5 * (zt @ zt) + torch.log(torch.diagonal(Lt)).sum()).item() + 0.5 * n * np.log(2 * np.pi)
    assert abs(v - ref) <= 1e-8 * abs(ref), (v, ref)
    # our own factor: L L^T = K on random probes
    info = handle.potrf_dev(K.data_ptr(), n, n)
    torch.cuda.synchronize()
    assert info == 0
    L = torch.tril(K)
    del K
    probe = torch.randn(n, 4, dtype=torch.float64, device='cuda')
    lhs = L @ (L.T @ probe)
    K2 = torch.empty((n, n), dtype=torch.float64, device='cuda')
    handle.kxx_dev(kh, K2.data_ptr())
    torch.cuda.synchronize()
    rhs = K2 @ probe
    assert ((lhs - rhs).abs().max() / rhs.abs().max()).item() < 1e-12
    assert ((L - Lt).abs().max()).item() < 1e-9
    # prediction against the torch solve
    fz, cov = handle.gpr_predict(kh, Z)
    Kxz = torch.from_numpy(gpr_oracle.kxz(lh, X, Z)).cuda()
    V = torch.linalg.solve_triangular(Lt, Kxz, upper=False)
    mref = (V.T @ zt).cpu().numpy()
    vref = (1.0 - (V * V).sum(0)).cpu().numpy()
    # 1e-9 relative to the scale of the posterior mean (elementwise ratios on the handful of
    # |mean| < 1e-3 entries only measure cancellation noise of both solvers)
    assert np.abs(fz - mref).max() / np.abs(mref).max() < 1e-9
    assert np.abs(cov - vref).max() < 1e-9


def test_fit_with_more_than_2_to_31_matrix_elements(handle):
    """Maximum sizes: N = 47104 (368 tiles) puts 2.2e9 elements = 17.7 GB into the factor, past every 32-bit element and
    byte offset.  Checked against cuSOLVER through torch (checker only) on the matrix our own kernel assembles."""
    import torch
    rng = np.random.default_rng(3)
    n, d = 47104, 8
    X = rng.random((n, d))
    w = rng.standard_normal(d)
    y = np.sin(X @ w) + 0.1 * rng.standard_normal(n)
    lh = np.log([0.5] * d + [1.0, 0.1])
    kh = khyp_of(lh)
    handle.set_train(X, y)
    v = handle.gpr_nlml(kh)
    v2 = handle.gpr_nlml(kh)
    assert v == v2                                            # run-to-run identical at this size too
    K = torch.empty((n, n), dtype=torch.float64, device='cuda')
    handle.kxx_dev(kh, K.data_ptr())
    torch.cuda.synchronize()
    tail = K[-200:, -200:].cpu().numpy()                      # the far corner of the matrix: offsets past 2^31 elements
    assert np.abs(tail - gpr_oracle.kxx(lh, X[-200:])).max() < 2e-14
    Lt = torch.linalg.cholesky(K)
    del K
    yt = torch.from_numpy(y).cuda()
    zt = torch.linalg.solve_triangular(Lt, yt[:, None], upper=False)[:, 0]
    ref = (0.5 * (zt @ zt) + torch.log(torch.diagonal(Lt)).sum()).item() + 0.5 * n * np.log(2 * np.pi)
    del Lt
    torch.cuda.empty_cache()
    assert abs(v - ref) <= 1e-8 * abs(ref), (v, ref)
    handle.set_train(X[:256], y[:256])                        # leave a small work space behind
    handle.gpr_nlml(kh)


# ---------------------------------------------------------------------------------------
# gradients (not in the reference: textbook formula, pinned by finite differences in test_oracle)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d", [(60, 3), (300, 1), (1000, 8), (1300, 2)])
def test_nlml_gradient_matches_oracle(handle, n, d):
    rng = np.random.default_rng(n + d)
    x = rng.random((n, d))
    y = np.sin(x @ rng.standard_normal(d)) + 0.1 * rng.standard_normal(n)
    lh = np.log(np.r_[0.4 + 0.4 * rng.random(d), 1.1, 0.15])
    handle.set_train(x, y)
    v, g = handle.gpr_nlml(khyp_of(lh), want_grad=True)
    ref = float(gpr_oracle.nlml(lh, x, y)[0, 0])
    gref = gpr_oracle.nlml_grad(lh, x, y)
    assert abs(v - ref) <= 1e-8 * abs(ref)
    assert np.abs(g - gref).max() <= 1e-8 * np.abs(gref).max(), (g, gref)
    assert abs(v - handle.gpr_nlml(khyp_of(lh))) <= 1e-13 * abs(v)


def test_batched_gradients_equal_single_calls(handle):
    rng = np.random.default_rng(12)
    n, d, B = 520, 2, 5
    x = 100 * rng.random((n, d))
    y = np.sin(x[:, 0] / 20) + 0.25 * rng.standard_normal(n)
    handle.set_train(x, y)
    lhs = np.log(np.c_[20 * np.exp(0.3 * rng.standard_normal((B, d))), np.sqrt(10) * np.ones(B), np.ones(B)])
    kh = np.array([khyp_of(l) for l in lhs])
    vals, grads, info = handle.gpr_nlml_batched(kh, want_grad=True)
    assert (info == 0).all()
    for b in range(B):
        v, g = handle.gpr_nlml(kh[b], want_grad=True)
        assert abs(v - vals[b]) <= 1e-13 * abs(v) and np.abs(g - grads[b]).max() <= 1e-11 * np.abs(g).max()
        gref = gpr_oracle.nlml_grad(lhs[b], x, y)
        assert np.abs(grads[b] - gref).max() <= 1e-8 * np.abs(gref).max()
