// tile_potrf.cu - factorisation of one 128x128 diagonal tile and its triangular inverse.
//
// Panel step of the blocked Cholesky (np.linalg.cholesky, GPr.py:62 / GPpref.py:128).  One CTA
// of 256 threads keeps the whole tile in REGISTERS (8x8 cells per thread, cell (r,c) with
// r = tr + 16a, c = tc + 16b); per column only the scaled column vector crosses shared memory.
//
// The inverse comes for free from the same right-looking sweep: factoring the augmented
// matrix [A; I] leaves L on top and I*L^-T = L^-T at the bottom.  L^-T is upper triangular, so
// its strict upper part is kept in the (otherwise unused) strict upper cells of the tile and its
// diagonal is 1/L_jj.  Step j:
//     d = sqrt(S[j][j]);  v[r] = S[r][j]/d (r != j),  v[j] = 1/d
//     S[r][c] -= v[r]*v[c]   for c > j and (r <= j  [rows of L^-T]  or  r >= c  [rows of L])
// With W = L^-1 the panel solve  X = P * L^-T  becomes the DMMA product  P * W^T.
//
// The 128 steps are a dependency chain, so the kernel is written for latency:
//   * every thread tracks the running diagonal entries of its own 8 columns (8 extra FMAs per
//     step), so the 16 owners of column j know S[j][j] without a broadcast;
//   * 1/d comes from rsqrt + one Newton correction (no sqrt -> divide chain);
//   * v is double buffered in shared memory: ONE barrier per column.
// Measured on B200 (tools/micro/lat.cu): a dependent DFMA/DMUL/DADD takes 23 clocks, rsqrt(double) 55, a
// 256-thread barrier 15 - the chain LDS -> FMA -> rsqrt -> correction -> scale -> STS -> barrier is ~230 clocks
// per column before any throughput term, and the kernel runs at ~745 clocks per column (50 us per tile).
// Variants that were measured slower and dropped: two software-pipelined forms (58.7 / 69.5 us), a blocked
// form with rank-16 updates of the later column blocks (54.5 us: fewer instructions, same chain); a cluster of two
// CTAs, one factoring and one building the inverse from columns pushed through distributed shared memory (75 us:
// the factor half alone is no faster than this kernel - halving the FMAs does not shorten the chain); a ninth
// "pivot warp" that runs the rsqrt chain of column j+1 under the update of column j (55 us: two named-barrier
// hand-offs per column cost more than the overlap wins).  All three were bitwise identical to this kernel.
// ncu (profiles/r01_ncu_tile_potrf.txt): 794 clocks per column, issue slots 28 % busy, FP64 pipe 22 %, stalls:
// barrier 32 %, fixed-latency wait 19 %, shared-memory scoreboard 12 % - a latency chain, not a throughput problem.
#include "gpb_kernels.cuh"

namespace gpb {

constexpr int TP_THREADS = 256;
constexpr int TP_PITCH = TILE + 1;
constexpr int TP_SMEM = TILE * TP_PITCH * 8;

// ---------------------------------------------------------------------------------------------------------
// Variant 2 (the default): the owners of column j+1 run its pivot chain in the MIDDLE of step j - right after the
// cells of their own column block have been updated and before the rest of the sweep - so that the rsqrt chain is
// issued under the FMAs of the other seven warps instead of after them.  Same cells, same operations per cell:
// bitwise identical to the kernel above, 46 instead of 50 us per tile (N = 1024: 0.565 vs 0.595 ms, N = 4096: 2.70 vs
// 2.81 ms).  On top of it a split barrier (the publishing warp only arrives, triple-buffered vector, three rotating
// barrier ids) was measured SLOWER again (0.626 ms at N = 1024) and dropped.  A probe that skips the publishing warp's
// remaining sweep altogether (wrong results, timing only) gains just 4 us per tile: deferring that work would not pay.
// With 512 threads (the 8 row groups of a thread split over two thread halves, four warps per scheduler) the tile takes
// 52 us: bitwise identical again, slower again (costlier barrier, the pivot computed twice, 128-register cap).
// A probe that drops every second barrier (wrong results, timing only) gains 4.5 us per tile: that bounds what a
// two-columns-per-barrier formulation (pair panels factored inside one warp, rank-2 sweeps) could win.
// ---------------------------------------------------------------------------------------------------------
template <int B>
__device__ __forceinline__ void upd_col_block(double (&c)[8][8], double (&dg)[8], const double (&vr)[8], const double (&vc)[8],
                                              const double (&vr_dg)[8], const double vr_inv, const int JB) {
  dg[B] = fma(-vc[B], vc[B], dg[B]);
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    if (a < JB) c[a][B] = fma(-vr[a], vc[B], c[a][B]);                 // inverse rows of earlier blocks
    else if (a == B) c[a][B] = fma(-vr_dg[a], vc[B], c[a][B]);         // diagonal 16x16 block
    else if (a > B) c[a][B] = fma(-vr[a], vc[B], c[a][B]);             // factor rows below
    else if (a == JB) c[a][B] = fma(-vr_inv, vc[B], c[a][B]);          // a == JB < B: inverse rows of this block
  }
}
template <int BN>
__device__ __forceinline__ void pivot_column(double (&c)[8][8], const double (&dg)[8], double* vb, int* fail, const int jn,
                                             const int tr) {
  const double ajj = dg[BN];
  if (tr == 0 && !(ajj > 0.0)) atomicMin(fail, jn);
  // 1/d and d each from the same rsqrt and ONE correction step of their own (<= 1 ulp), computed side by side:
  // 1/d = r0 + (r0/2)(1 - a r0^2) is ready two dependent operations earlier than through d (it sits on the chain,
  // d does not - it is only stored)
  const double r0 = rsqrt(ajj);
  const double d0 = ajj * r0;
  const double invd = fma(0.5 * r0, fma(-d0, r0, 1.0), r0);
  const double d = fma(0.5 * r0, fma(-d0, d0, ajj), d0);
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int r = tr + 16 * a;
    if (r != jn) {
      c[a][BN] *= invd;
      vb[r] = c[a][BN];
    } else {
      c[a][BN] = d;
      vb[r] = invd;
    }
  }
}
template <int JB, bool LAST>
__device__ __forceinline__ void v2_step(double (&c)[8][8], double (&dg)[8], double (*v)[TILE], int* fail, const int jj,
                                        const int tr, const int tc, const bool ge) {
  const int j = 16 * JB + jj;
  const double* vb = v[j & 1];
  double vr[8], vc[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) vr[a] = vb[tr + 16 * a];
#pragma unroll
  for (int b = JB; b < 8; ++b) vc[b] = vb[tc + 16 * b];
  const bool p_row = (tr <= jj);
  vc[JB] = (tc > jj) ? vc[JB] : 0.0;
  const double vr_inv = p_row ? vr[JB] : 0.0;
  double vr_dg[8];
#pragma unroll
  for (int a = JB; a < 8; ++a) vr_dg[a] = (ge || (a == JB && p_row)) ? vr[a] : 0.0;
  constexpr int BN = LAST ? JB + 1 : JB;                 // column block of column j+1
  upd_col_block<JB>(c, dg, vr, vc, vr_dg, vr_inv, JB);
  if (LAST && BN < 8) upd_col_block<(BN < 8 ? BN : 7)>(c, dg, vr, vc, vr_dg, vr_inv, JB);
  if (BN < 8) {
    const int tcn = LAST ? 0 : jj + 1;
    if (tc == tcn) pivot_column<(BN < 8 ? BN : 7)>(c, dg, v[(j + 1) & 1], fail, j + 1, tr);
  }
#pragma unroll
  for (int b = BN + 1; b < 8; ++b) {
    dg[b] = fma(-vc[b], vc[b], dg[b]);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      if (a < JB) c[a][b] = fma(-vr[a], vc[b], c[a][b]);
      else if (a == b) c[a][b] = fma(-vr_dg[a], vc[b], c[a][b]);
      else if (a > b) c[a][b] = fma(-vr[a], vc[b], c[a][b]);
      else if (a == JB) c[a][b] = fma(-vr_inv, vc[b], c[a][b]);
    }
  }
  __syncthreads();
}
template <int JB>
__device__ __forceinline__ void v2_block(double (&c)[8][8], double (&dg)[8], double (*v)[TILE], int* fail, const int tr,
                                         const int tc) {
  const bool ge = (tr >= tc);
  for (int jj = 0; jj < 15; ++jj) v2_step<JB, false>(c, dg, v, fail, jj, tr, tc, ge);
  v2_step<JB, true>(c, dg, v, fail, 15, tr, tc, ge);
}

__global__ void __launch_bounds__(TP_THREADS, 1) tile_potrf_inv_kernel2(const TilePotrfArgs p) {
  extern __shared__ double S[];
  __shared__ double v[2][TILE];
  __shared__ int fail;
  const int t = threadIdx.x;
  const int tc = t >> 4, tr = t & 15;
  const int batch = blockIdx.x;
  double* Ab = p.A + batch * p.a_batch_stride + static_cast<int64_t>(p.k) * TILE * p.lda + p.k * TILE;
  pdl_trigger();
  if (t == 0) fail = TILE;
  pdl_wait();
  double c[8][8], dg[8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int r = tr + 16 * a, cc = tc + 16 * b;
      c[a][b] = (r >= cc) ? Ab[static_cast<int64_t>(r) * p.lda + cc] : 0.0;
    }
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    const int cc = tc + 16 * b;
    dg[b] = Ab[static_cast<int64_t>(cc) * p.lda + cc];
  }
  __syncthreads();
  if (tc == 0) pivot_column<0>(c, dg, v[0], &fail, 0, tr);
  __syncthreads();
  v2_block<0>(c, dg, v, &fail, tr, tc);
  v2_block<1>(c, dg, v, &fail, tr, tc);
  v2_block<2>(c, dg, v, &fail, tr, tc);
  v2_block<3>(c, dg, v, &fail, tr, tc);
  v2_block<4>(c, dg, v, &fail, tr, tc);
  v2_block<5>(c, dg, v, &fail, tr, tc);
  v2_block<6>(c, dg, v, &fail, tr, tc);
  v2_block<7>(c, dg, v, &fail, tr, tc);
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) S[(tr + 16 * a) * TP_PITCH + tc + 16 * b] = c[a][b];
  __syncthreads();
  double* Dk = p.Dinv + batch * p.d_batch_stride + static_cast<int64_t>(p.k) * TILE * TILE;
  for (int idx = t; idx < TILE * TILE; idx += TP_THREADS) {
    const int r = idx >> 7, cc = idx & 127;
    if (cc <= r) Ab[static_cast<int64_t>(r) * p.lda + cc] = S[r * TP_PITCH + cc];
    double w = 0.0;
    if (cc < r) w = S[cc * TP_PITCH + r];
    else if (cc == r) w = 1.0 / S[r * TP_PITCH + r];
    Dk[idx] = w;
  }
  if (t < TILE) p.diag[batch * p.diag_batch_stride + p.k * TILE + t] = S[t * TP_PITCH + t];
  if (t == 0 && fail < TILE) atomicCAS(p.info + batch, 0, p.k * TILE + fail + 1);
}

// variant 3 (default): blocked inside the CTA, warp-shuffle base case + DMMA block products (tile_potrf3.cu)
// variant 2          : the register-resident sweep above (kept as a cross-check: tests/test_gpu_edges.py)
void tile_potrf3_init();
void launch_tile_potrf3(const TilePotrfArgs& a, int batch, cudaStream_t st, bool pdl, bool refine);

static int g_potrf_variant = 3;
static int g_potrf_refine = 1;        // variant 3: one correction step on rsqrt for 1/d and d (<= 1 ulp) or the bare rsqrt
void tile_potrf_set_variant(int v) { g_potrf_variant = (v == 2) ? 2 : 3; }
void tile_potrf_set_refine(int on) { g_potrf_refine = on != 0; }
bool tile_potrf_fuses_rhs() { return g_potrf_variant != 2; }

void tile_potrf_init() {
  GPB_CUDA(cudaFuncSetAttribute(tile_potrf_inv_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
  tile_potrf3_init();
}

void launch_tile_potrf_inv(TilePotrfArgs a, int batch, cudaStream_t st) {
  const bool pdl = g_pdl != 0 && a.pdl != 0;
  if (g_potrf_variant == 2) {
    launch_chain(tile_potrf_inv_kernel2, dim3(batch), dim3(TP_THREADS), TP_SMEM, st, pdl, a);
    return;
  }
  launch_tile_potrf3(a, batch, st, pdl, g_potrf_refine != 0);
}

}  // namespace gpb
