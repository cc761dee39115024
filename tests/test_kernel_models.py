"""CPU models of the CUDA kernels' index arithmetic (no GPU needed).

These tests re-state, in numpy, the exact addressing and scheduling formulas used by the
kernels in gptest_b200/csrc and check them against plain linear algebra:

  * dmma_gemm.cu : TMA 128B-swizzle placement, "parity" fragment addressing, DMMA.8x8x4 lane
                   layout and the epilogue's (lane -> 4 consecutive columns) mapping;
  * tile_potrf.cu: the augmented right-looking sweep that leaves L and L^-T in one tile;
  * chol.cu      : the blocked sweep with appended right-hand-side rows and the tile decode of
                   the trapezoid regions.

They are how the kernels were derived; they stay as regression tests for the formulas.
"""
import numpy as np
import pytest


# ---------------------------------------------------------------------------------------
def swizzle128_offset(r, c):
    """byte offset of element (row r, col c) of a 128-row x 16-double box under SWIZZLE_128B"""
    return r * 128 + (((c >> 1) ^ (r & 7)) << 4) + (c & 1) * 8


def tma_box(mat):
    """what TMA leaves in shared memory for a (128,16) fp64 box: flat array of 2048 doubles"""
    sm = np.full(128 * 16, np.nan)
    for r in range(128):
        for c in range(16):
            off = swizzle128_offset(r, c)
            assert off % 8 == 0
            sm[off // 8] = mat[r, c]
    assert not np.isnan(sm).any()          # the swizzle is a bijection
    return sm


def test_fragment_addressing_and_dmma_layout():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((128, 16))
    B = rng.standard_normal((128, 16))
    sa, sb = tma_box(A), tma_box(B)
    C = np.zeros((128, 128))
    for warp in range(8):
        wm, wn = warp >> 2, warp & 3
        acc = np.zeros((32, 8, 4, 2))                    # lane, mi, nj, e
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            th = t >> 1
            a_off, b_off, xr = [0, 0], [0, 0], [0, 0]
            for par in range(2):
                xr[par] = ((g & 3) << 1) | par
                a_off[par] = (wm * 64 + 2 * g + par) * 128 + (t & 1) * 8
                b_off[par] = (wn * 32 + 2 * g + par) * 128 + (t & 1) * 8
            for kk in range(4):
                af, bf = np.zeros(8), np.zeros(4)
                for par in range(2):
                    chunk = ((2 * kk + th) ^ xr[par]) << 4
                    for grp in range(4):
                        af[grp * 2 + par] = sa[(a_off[par] + grp * 2048 + chunk) // 8]
                    for grp in range(2):
                        bf[grp * 2 + par] = sb[(b_off[par] + grp * 2048 + chunk) // 8]
                # stash per-lane fragments for the warp-wide mma below
                if kk == 0:
                    pass
                acc_lane = (af, bf)
                # emulate mma.m8n8k4 across the warp: needs all lanes -> do it lane-major afterwards
                test_fragment_addressing_and_dmma_layout.frag[(warp, lane, kk)] = acc_lane
        # warp-wide DMMA: D[g][2t+e] += sum_k A[g][k] B[k][2t+e]; lane l holds A[l/4][l%4], B[l%4][l/4]
        for kk in range(4):
            for mi in range(8):
                for nj in range(4):
                    Af = np.zeros((8, 4))
                    Bf = np.zeros((4, 8))
                    for lane in range(32):
                        af, bf = test_fragment_addressing_and_dmma_layout.frag[(warp, lane, kk)]
                        Af[lane >> 2, lane & 3] = af[mi]
                        Bf[lane & 3, lane >> 2] = bf[nj]
                    D = Af @ Bf
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        acc[lane, mi, nj, 0] += D[g, 2 * t]
                        acc[lane, mi, nj, 1] += D[g, 2 * t + 1]
        # epilogue mapping
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for gm in range(4):
                for pm in range(2):
                    row = wm * 64 + 16 * gm + 2 * g + pm
                    mi = gm * 2 + pm
                    for gn in range(2):
                        col = wn * 32 + 16 * gn + 4 * t
                        C[row, col + 0] = acc[lane, mi, gn * 2, 0]
                        C[row, col + 1] = acc[lane, mi, gn * 2 + 1, 0]
                        C[row, col + 2] = acc[lane, mi, gn * 2, 1]
                        C[row, col + 3] = acc[lane, mi, gn * 2 + 1, 1]
    assert np.allclose(C, A @ B.T, rtol=0, atol=1e-12)


test_fragment_addressing_and_dmma_layout.frag = {}


def test_fragment_loads_are_bank_conflict_free():
    """each half warp of a 64-bit fragment load must touch all 32 banks exactly once"""
    for wm in range(2):
        for par in range(2):
            for kk in range(4):
                for grp in range(4):
                    for half in range(2):
                        banks = []
                        for lane in range(16 * half, 16 * half + 16):
                            g, t = lane >> 2, lane & 3
                            xr = ((g & 3) << 1) | par
                            off = (wm * 64 + 2 * g + par) * 128 + (t & 1) * 8 + grp * 2048 + ((((2 * kk) + (t >> 1)) ^ xr) << 4)
                            banks += [(off // 4) % 32, (off // 4 + 1) % 32]
                        assert sorted(banks) == list(range(32))


# ---------------------------------------------------------------------------------------
def tile_potrf_inv_model(A):
    """tile_potrf.cu step by step: returns (L, W = L^-1) from one in-place sweep"""
    n = A.shape[0]
    S = np.tril(A).copy()                      # strict upper cells start at 0 (identity's off-diagonal)
    for j in range(n):
        d = np.sqrt(S[j, j])
        S[j, j] = d
        v = S[:, j] / d
        v[j] = 1.0 / d
        S[:, j] = np.where(np.arange(n) != j, v, S[:, j])
        for c in range(j + 1, n):
            rows = np.r_[0:j + 1, c:n]
            S[rows, c] -= v[rows] * v[c]
    L = np.tril(S)
    W = np.triu(S, 1).T + np.diag(1.0 / np.diag(S))
    return L, W


def test_tile_potrf_overlay_trick():
    rng = np.random.default_rng(1)
    n = 24
    M = rng.standard_normal((n, n))
    A = M @ M.T + n * np.eye(n)
    L, W = tile_potrf_inv_model(A)
    Lref = np.linalg.cholesky(A)
    assert np.allclose(L, Lref, atol=1e-12)
    assert np.allclose(W, np.linalg.inv(Lref), atol=1e-12)


# ---------------------------------------------------------------------------------------
def decode_tri(idx, j0, R, i_off):
    """dmma_gemm.cu decode_tile for tri == 1"""
    H = R - j0 - i_off
    b = 2.0 * H + 1.0
    c = int((b - np.sqrt(b * b - 8.0 * idx)) * 0.5)
    c = max(c, 0)
    while (c + 1) * H - (c + 1) * c // 2 <= idx:
        c += 1
    while c * H - c * (c - 1) // 2 > idx:
        c -= 1
    jt = j0 + c
    return jt + i_off + (idx - (c * H - c * (c - 1) // 2)), jt


@pytest.mark.parametrize("j0,j1,R,i_off", [(0, 1, 5, 1), (2, 7, 9, 0), (0, 128, 129, 0), (3, 4, 4, 0), (5, 7, 130, 0)])
def test_trapezoid_decode(j0, j1, R, i_off):
    ncols = j1 - j0
    H = R - j0 - i_off
    ntiles = ncols * H - ncols * (ncols - 1) // 2
    got = [decode_tri(i, j0, R, i_off) for i in range(ntiles)]
    want = [(i, j) for j in range(j0, j1) for i in range(j + i_off, R)]
    assert got == want


def blocked_sweep_model(A, extra, T, nb, lookahead_order=False):
    """chol.cu: blocked right-looking sweep on tiles of size T with appended rows `extra`.

    Returns (L, extra @ L^-T).  Uses exactly the panel / trail decomposition of the driver.
    """
    n = A.shape[0]
    M = np.vstack([np.tril(A), extra]).copy()
    nt = n // T
    Rrows = M.shape[0]

    def trsm(k):
        Lkk = np.tril(M[k * T:(k + 1) * T, k * T:(k + 1) * T])
        W = np.linalg.inv(Lkk)
        M[(k + 1) * T:Rrows, k * T:(k + 1) * T] = M[(k + 1) * T:Rrows, k * T:(k + 1) * T] @ W.T

    def update(c0, c1, ka, kb):
        for j in range(c0, c1):
            P_j = M[j * T:(j + 1) * T, ka * T:kb * T]
            M[j * T:Rrows, j * T:(j + 1) * T] -= M[j * T:Rrows, ka * T:kb * T] @ P_j.T

    def panel(kb, kend):
        for k in range(kb, kend):
            blk = M[k * T:(k + 1) * T, k * T:(k + 1) * T]
            blk[:] = np.linalg.cholesky(np.tril(blk) + np.tril(blk, -1).T)
            trsm(k)
            update(k + 1, kend, k, k + 1)

    if not lookahead_order:
        for kb in range(0, nt, nb):
            kend = min(kb + nb, nt)
            panel(kb, kend)
            update(kend, nt, kb, kend)
    else:
        panel(0, min(nb, nt))
        for kb in range(0, nt, nb):
            kend = min(kb + nb, nt)
            if kend >= nt:
                break
            nend = min(kend + nb, nt)
            update(kend, nend, kb, kend)
            panel(kend, nend)
            update(nend, nt, kb, kend)
    return np.tril(M[:n]), M[n:]


@pytest.mark.parametrize("nb,la", [(1, False), (2, False), (2, True), (3, True), (4, True)])
def test_blocked_sweep_with_appended_rows(nb, la):
    rng = np.random.default_rng(2)
    T, nt = 8, 7
    n = T * nt
    Mx = rng.standard_normal((n, n))
    A = Mx @ Mx.T + n * np.eye(n)
    extra = rng.standard_normal((11, n))
    L, XT = blocked_sweep_model(A, extra, T, nb, la)
    Lref = np.linalg.cholesky(A)
    assert np.allclose(L, Lref, atol=1e-10)
    assert np.allclose(XT, np.linalg.solve(Lref, extra.T).T, atol=1e-10)
