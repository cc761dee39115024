// gpb_kernels.cuh - launch interfaces of the sm_100a kernels of libgpb200.
#pragma once
#include "gpb_common.cuh"

namespace gpb {

// ---------------------------------------------------------------------------------------
// TMA-fed FP64 DMMA "NT" tile kernel:  C(it,jt) (op)= sum_k A(it, k) * B(jt, k)^T
// on 128x128 tiles of row-major matrices.  One CTA per tile; blockIdx.y = batch entry.
// ---------------------------------------------------------------------------------------
struct GemmArgs {
  double* C;                // output matrix base
  int64_t ldc;              // leading dimension of C (elements, even)
  int64_t c_batch_stride;   // elements between batch entries of C
  int rows_total;           // rows >= rows_total are never stored
  // tile region.  tri == 1: columns jt in [j0, j1), rows it in [jt + i_off, R) (a trapezoid,
  // enumerated column by column).  tri == 0: rows it in [i0, R) for every column.
  int j0, j1, R, tri, i_off, i0;
  // All tile coordinates (j0, j1, R, i_off, i0, it, jt) are in units of the kernel's tile edge T
  // (128 or 64); row / k offsets are in elements.
  // operands.  A slab s: mapA box at (ka0 + 16 s, a_row0 + T it, batch);
  //            B slab s: mapB box at (kb0 + 16 s, b_row0 + T jt, batch).
  // k_from_row != 0: the k range of tile (it, jt) starts at column T*it (operands that are
  // upper triangular: U U^T products), i.e. ka0 = kb0 = T it and nk = (k_end - T it) / 16.
  int ka0, kb0, nk, a_row0, b_row0, k_from_row, k_end;
  int epi;                  // 0: C = acc     1: C = C - acc
  int ntiles;               // region tiles per batch entry (filled by the launcher)
  int nbatch;               // batch entries (filled by the launcher); work list = nbatch x tiles
  int num_sms;              // filled by the launcher
  long long stagger_clk;    // > 0: delay (SM clocks) of the second resident CTA per SM in the first wave
  int pdl;                  // launch with programmatic stream serialisation if the launch is small (see gpb_common.cuh)
  int no_stagger;           // chunked schedule: CTAs of several launches share the SMs and dephase by themselves
  int b_tri;                // the B tile is lower triangular (an inverted diagonal tile, the panel TRSM): a warp skips the
                            // k slabs right of its last output column (they multiply zeros)
  // Forward substitution riding on the panel TRSM: with z = the solved tile z_k of the right-hand side, the CTA that
  // has produced X = rows of L[., tile k] takes  X[row][:] . z  out of the running right-hand side of its rows (it owns
  // all 128 columns of them), so L is never read again for the solve.  The right-hand side is kept as EIGHT partial
  // vectors, one per 16-column group of the tile: r[row] = sum_g r_g[row] (summed in order by the reader, the
  // diagonal-tile kernel).  Group g of a row is owned by exactly one lane of one CTA of the launch, so there is no
  // reduction across warps and no barrier in the epilogue, and the order of every sum is fixed.  Null: off.
  const double* gemv_z;     // per batch entry: 128 doubles (z_k)
  double* gemv_r;           // per batch entry: 8 partial right-hand sides, group g at + g * gemv_gs, indexed by global row
  int64_t gemv_bs;          // elements between batch entries of r
  int64_t gemv_gs;          // elements between the partial vectors of one batch entry
  int64_t gemv_zbs;         // elements between batch entries of z
  int sym_lower;            // symmetric update of a factorisation: only the lower triangle of C is ever read, so warp
                            // tiles that lie strictly above the diagonal are neither computed nor stored
};
int gemm_region_tiles(const GemmArgs& a);       // host: number of tiles of the region
// one tensor map per tile edge: the TMA box height is part of the map
struct TileMaps {
  CUtensorMap m128, m64, m32;
  const CUtensorMap& get(int tile) const { return tile == 128 ? m128 : (tile == 64 ? m64 : m32); }
};
void launch_dmma_gemm(const CUtensorMap& mapA, const CUtensorMap& mapB, GemmArgs a, int batch,
                      cudaStream_t st, int tile);
// region given in 128-tile units -> the same region in 64-tile units
GemmArgs gemm_args_to_64(const GemmArgs& a);
void dmma_gemm_init();                           // sets the dynamic smem attribute once
void dmma_gemm_set_persistent(int waves);        // 0 (default): one CTA per tile; n: persistent grid of n waves
void dmma_gemm_set_pdl(int mode);               // programmatic dependent launch: 0 off, 1 small launches (default), 2 all
void dmma_gemm_set_trsm_balance(int mode);     // 64 x 128 panel TRSM with a triangular B: 1 (default) 8-warp CTAs, 0 by launch size as elsewhere
void dmma_gemm_set_trsm_persist(int waves);    // > 0: the panel TRSM runs as that many resident waves of CTAs walking the tile list
void dmma_gemm_set_fine_warps(int on);          // 1 (default): 8-warp CTAs for launches of at most one CTA per SM
void dmma_gemm_set_stagger(int on);              // 1 (default): phase-shift co-resident CTAs of the 2-per-SM variants

// ---------------------------------------------------------------------------------------
// Diagonal-tile factorisation: L = chol(A_kk) in place (lower), W = L^-1 to Dinv[k] (dense
// 128x128, upper part zero), diag(L) to diag[k*128..], first failing pivot to *info.
// ---------------------------------------------------------------------------------------
struct TilePotrfArgs {
  double* A;
  int64_t lda, a_batch_stride;
  int k;                    // tile index on the diagonal
  double* Dinv;             // per batch: nt*128 rows x 128 cols
  int64_t d_batch_stride;
  double* diag;             // per batch: n_pad doubles
  int64_t diag_batch_stride;
  int* info;                // per batch: first failing pivot (1-based global index), 0 = ok
  int pdl;                  // launch with programmatic stream serialisation
  // optional right-hand side riding on the factorisation (GemmArgs::gemv_*): the kernel also solves tile k of it,
  // z_k = W_k r_k, once W is complete (variant 3 only; see tile_potrf_fuses_rhs)
  const double* rhs_r;      // 8 partial vectors per batch entry (see GemmArgs::gemv_r)
  double* rhs_z;
  int64_t rhs_bs, rhs_gs, rhs_zbs;
  long long* dbg;           // optional (tools/micro/tp3_bench.cu): clock64() of the phase boundaries, per warp; normally null
};
void launch_tile_potrf_inv(TilePotrfArgs a, int batch, cudaStream_t st);
void tile_potrf_init();
void tile_potrf_set_variant(int v);          // 3 (default): blocked in-CTA kernel (tile_potrf3.cu)  2: register-resident sweep
void tile_potrf_set_refine(int on);
bool tile_potrf_fuses_rhs();                 // the selected variant computes z_k = W_k r_k itself (TilePotrfArgs::rhs_*)          // variant 3: corrected (default) or bare rsqrt pivots

// ---------------------------------------------------------------------------------------
// SE-ARD covariance assembly
// ---------------------------------------------------------------------------------------
// XsT[d][i] = X[i][d] / ell[d] (d-major, pitch ld_t), sq[i] = sum_d Xs^2 (same FMA chain as the
// tile kernel's dot product, so the expanded-form distance of a point to itself is exactly 0).
void launch_se_prep(const double* X, int64_t n, int d, const double* ell_dev, double* XsT,
                    int64_t ld_t, double* sq, int batch, int64_t ell_batch_stride,
                    int64_t xs_batch_stride, int64_t sq_batch_stride, cudaStream_t st);
struct SeArgs {
  const double* rT;  int64_t r_ld;  const double* r_sq;  int64_t n_rows_valid;  // row points
  const double* cT;  int64_t c_ld;  const double* c_sq;  int64_t n_cols_valid;  // column points
  int64_t xs_batch_stride, sq_batch_stride;      // per-batch strides of the scaled points
  int d;
  double* out; int64_t ld; int64_t out_batch_stride;
  int64_t rows_pad, cols_pad;     // extent written (multiples of 64); beyond *_valid: identity / 0
  const double* hyp_dev;          // per batch: [sf2, sn2]
  int mode;                       // 0: symmetric, write both triangles  1: symmetric, lower tiles only
                                  // 2: rectangular (no noise term, pad = 0)
                                  // 3: rectangular, raw squared distances (no exp)
  int clip;                       // 1: clamp r^2 at 0 (GPy RBF semantics)
  int kind;                       // radial function: 0 squared exponential, 1 Matern 3/2, 2 Matern 5/2 (gpb_exp.cuh)
};
void launch_se_build(const SeArgs& a, int batch, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// finishing reductions (deterministic: fixed-order tree, no atomics)
// ---------------------------------------------------------------------------------------
// out[b] = 0.5*sum(z^2) + sum(log diag) + 0.5*n_valid*log(2 pi)
void launch_nlml_finish(const double* z, int64_t z_batch_stride, const double* diag,
                        int64_t diag_batch_stride, int64_t n_pad, int64_t n_valid, double* out,
                        int batch, cudaStream_t st);
// mean[i] = VT[i,:] . z ; var[i] = sf2 - |VT[i,:]|^2   (VT rows have pitch ld)
void launch_predict_finish(const double* VT, int64_t ld, const double* z, int64_t n_pad, int64_t m,
                           const double* hyp_dev, double* mean, double* var, cudaStream_t st);
// generic helpers
void launch_set_y_rows(double* A, int64_t batch_stride, int64_t row_off, const double* y, int64_t n,
                       int64_t n_pad, double mean, int batch, cudaStream_t st);
void launch_row_dot(const double* M, int64_t ld, int64_t m_bs, const double* x, int64_t x_bs, int64_t nrows,
                    int64_t ncols, int upper, double* out, int64_t out_bs, int batch, cudaStream_t st);
// x <- L^-1 r per batch entry by blocked forward substitution with the inverted diagonal tiles (r: scratch)
void launch_trsv_l(const double* L, int64_t ld, int64_t l_bs, const double* Dinv, int64_t d_bs, int64_t n_pad,
                   double* r, int64_t r_bs, double* x, int64_t x_bs, int batch, cudaStream_t st, int k_begin = 0);
// the same for m <= 16 right-hand sides stored as rows with pitch `pitch` that share one L (read once per step)
void launch_trsv_l_multi(const double* L, int64_t ld, const double* Dinv, int64_t n_pad, double* r, double* x,
                         int64_t pitch, int m, cudaStream_t st);
void launch_fill(double* p, int64_t n, double v, cudaStream_t st);
void launch_copy_sub_mean(double* dst, const double* src, int64_t n, double mean, cudaStream_t st);

// micro-benchmarks (kind 0: DMMA, 1: DFMA): returns flops executed; caller times it
double launch_microbench(int kind, double* sink, cudaStream_t st);

}  // namespace gpb
