// chol.cu - host driver of the blocked right-looking Cholesky sweep.
//
// Replaces np.linalg.cholesky + the two np.linalg.solve calls of GPr.py:62-63 / GPpref.py:128-129.
//
// The matrix is cut into 128x128 tiles; an outer block column is nb_tiles tiles wide.
//   panel(kb):   for each tile column k of the block
//                   tile_potrf_inv(k,k)           L_kk and W_k = L_kk^-1           (1 CTA)
//                   A[i,k] <- A[i,k] W_k^T        all rows below, DMMA product     (R-k-1 CTAs)
//                   A[i,j] -= A[i,k] A[j,k]^T     remaining columns of the block   (K = 128)
//   trail(kb):   A[i,j] -= sum_k A[i,k] A[j,k]^T  everything right of the block    (K = 128 nb)
// Rows appended under the symmetric part are right-hand sides stored as rows: they receive the
// same TRSM/update and end up multiplied by L^-T, i.e. the forward solves L^-1 y (and
// L^-1 Kxz for the predictive variance) cost no extra pass.
//
// Look-ahead: the columns of the next panel are updated first, then the next panel is factored
// on a high-priority side stream while the main stream updates the rest of the trailing matrix.
// Both streams are ordered with events only - the host never blocks inside the sweep.
#include <cmath>

#include "gpb_context.cuh"

namespace gpb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    GPB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) throw Error{"cuTensorMapEncodeTiled not available in this driver"};
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 3-D fp64 tensor map {cols, rows, batch}; box = {16 doubles (128 B), 128 rows, 1}; 128B swizzle;
// out-of-bounds rows read as zero (partial last row tile of the appended rows).
void make_tensor_map(CUtensorMap* map, const double* base, int64_t cols, int64_t rows, int64_t batch,
                     int64_t row_pitch, int64_t batch_pitch, int box_rows) {
  GPB_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "matrix base must be 16-byte aligned");
  GPB_REQUIRE(row_pitch % 2 == 0, "leading dimension must be even (16-byte row pitch)");
  if (batch_pitch <= 0) batch_pitch = rows * row_pitch;
  GPB_REQUIRE(batch_pitch % 2 == 0, "batch pitch must be even");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(batch < 1 ? 1 : batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_pitch) * 8, static_cast<cuuint64_t>(batch_pitch) * 8};
  cuuint32_t box[3] = {GEMM_KB, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides,
                           box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error{"cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r))};
}

void make_tile_maps(TileMaps* maps, const double* base, int64_t cols, int64_t rows, int64_t batch,
                    int64_t row_pitch, int64_t batch_pitch) {
  make_tensor_map(&maps->m128, base, cols, rows, batch, row_pitch, batch_pitch, 128);
  make_tensor_map(&maps->m64, base, cols, rows, batch, row_pitch, batch_pitch, 64);
  make_tensor_map(&maps->m32, base, cols, rows, batch, row_pitch, batch_pitch, 32);
}

void finalize_factor_mat(FactorMat& m) {
  GPB_REQUIRE(m.n_pad % TILE == 0, "n_pad must be a multiple of 128");
  make_tile_maps(&m.mapA, m.A, m.ld, m.rows_total, m.batch, m.ld, m.batch_stride);
  make_tile_maps(&m.mapD, m.Dinv, TILE, m.n_pad, m.batch, TILE, m.dinv_bs);
}

GemmArgs gemm_args_to_64(const GemmArgs& a) {
  GemmArgs b = a;
  b.j0 = 2 * a.j0; b.j1 = 2 * a.j1; b.i_off = 2 * a.i_off; b.i0 = 2 * a.i0;
  const int r64 = (a.rows_total + 63) / 64;
  b.R = 2 * a.R < r64 ? 2 * a.R : r64;
  return b;
}

namespace {

struct Sweep {
  gpb_handle* h;
  FactorMat& m;
  bool factor;
  int nt, R;
  int x0 = 0, xn = 0;     // non-factor mode: tile rows [x0, x0 + xn) are swept
  bool grow = false;      // identity right-hand sides: tile row x0 + q is zero left of column q
  bool pdl = false;       // small launches of this sweep chain with programmatic dependent launch
  bool big_tiles = false; // chunked schedule: the launch is small but shares the machine with others - keep the big tiles

  // last tile row (exclusive) that can be non-zero once tile column k has been processed
  int r_at(int k) const {
    if (factor) return R;
    if (!grow) return x0 + xn;
    return x0 + (k + 1 < xn ? k + 1 : xn);
  }
  GemmArgs base(int k) const {
    GemmArgs a{};
    a.C = m.A;
    a.ldc = m.ld;
    a.c_batch_stride = m.batch_stride;
    a.rows_total = static_cast<int>(m.rows_total);
    a.R = r_at(k);
    a.pdl = pdl;
    a.no_stagger = big_tiles;
    if (factor) {
      a.tri = 1;
    } else {
      a.tri = 0;
      a.i0 = x0;
    }
    return a;
  }
  // 64-tiles when the launch cannot fill the machine with 128-tiles (panel critical path, tail)
  void launch(const GemmArgs& a, const TileMaps& ma, const TileMaps& mb, cudaStream_t st) {
    const int n128 = gemm_region_tiles(a);
    if (n128 <= 0) return;
    if (!big_tiles && a.tri && a.j1 - a.j0 == 1 && static_cast<int64_t>(n128) * m.batch <= h->thin_tile_max) {
      // one tile column with few tiles: the update of the next panel's column, i.e. the critical chain of a small
      // matrix.  A lone 64x64 CTA-tile needs 4.2 us of its SM's FP64 pipe for K = 128; 32 x 64 tiles put the same work
      // on four times as many SMs (rectangular region: the diagonal tile is computed whole, its upper half is unused).
      GemmArgs b = a;
      b.tri = 0;
      b.i0 = 4 * (a.j0 + a.i_off);
      const int r32 = (a.rows_total + 31) / 32;
      b.R = 4 * a.R < r32 ? 4 * a.R : r32;
      b.j0 = 2 * a.j0; b.j1 = 2 * a.j1; b.i_off = 0;
      launch_dmma_gemm(ma.m32, mb.m64, b, m.batch, st, 3264);
      ++h->launches;
      return;
    }
    // batches of matrices with short k (the plain-order sweeps, K <= 512): 64 x 64 CTA-tiles, three per SM, whatever the
    // tile count - finer skipping of the unused half of diagonal tiles and more CTAs to cover the fill and drain of
    // an 8..32-slab k loop (1024 x N=2048: 107.2 -> 103.4 ms, 128 problems 13.81 -> 13.40 ms)
    const bool short_k_batch = m.batch > 1 && !a.k_from_row && a.nk * GEMM_KB <= h->batch_small_k;
    if (!big_tiles && (static_cast<int64_t>(n128) * m.batch < h->small_tile_threshold || short_k_batch)) {
      launch_dmma_gemm(ma.m64, mb.m64, gemm_args_to_64(a), m.batch, st, 64);
    } else if (h->split_tiles) {
      launch_dmma_gemm(ma.m128, mb.m64, a, m.batch, st, 12864);
    } else {
      launch_dmma_gemm(ma.m128, mb.m128, a, m.batch, st, 128);
    }
    ++h->launches;
  }
  // A[i,k] <- A[i,k] * W_k^T for the rows below tile (k,k) (and the appended rows).  In place: a CTA
  // must own all 128 output columns of its rows (they are also its A operand), so the small-tile
  // variant is 64 rows x 128 columns.
  void trsm(int k, cudaStream_t st) {
    GemmArgs a = base(k);
    a.tri = 0;                                  // one tile column: rows [first, R)
    a.i0 = factor ? k + 1 : x0;
    a.j0 = k; a.j1 = k + 1;
    a.ka0 = k * TILE; a.kb0 = 0; a.nk = TILE / GEMM_KB; a.b_row0 = 0;
    a.epi = 0;
    a.b_tri = h->tri_skip;
    if (factor && m.rhs_r) {
      a.gemv_z = m.rhs_z + static_cast<int64_t>(k) * TILE;
      a.gemv_r = m.rhs_r;
      a.gemv_bs = m.rhs_bs;
      a.gemv_gs = m.rhs_gs;
      a.gemv_zbs = m.rhs_zbs;
    }
    const int n128 = gemm_region_tiles(a);
    if (n128 <= 0) return;
    if (static_cast<int64_t>(n128) * m.batch <= h->thin_tile_max) {
      // few rows (the panel of a small matrix, on the critical chain): 32-row CTA-tiles, four per 128-row tile
      GemmArgs b = a;
      b.i0 = 4 * a.i0;
      const int r32 = (a.rows_total + 31) / 32;
      b.R = 4 * a.R < r32 ? 4 * a.R : r32;
      launch_dmma_gemm(m.mapA.m32, m.mapD.m128, b, m.batch, st, 32128);
      ++h->launches;
      return;
    }
    if (static_cast<int64_t>(n128) * m.batch < h->trsm_tile_threshold) {
      GemmArgs b = a;
      b.i0 = 2 * a.i0;
      const int r64 = (a.rows_total + 63) / 64;
      b.R = 2 * a.R < r64 ? 2 * a.R : r64;
      launch_dmma_gemm(m.mapA.m64, m.mapD.m128, b, m.batch, st, 64128);
    } else {
      launch_dmma_gemm(m.mapA.m128, m.mapD.m128, a, m.batch, st, 128);
    }
    ++h->launches;
  }
  // A[i,j] -= sum_{k in [ka,kb)} A[i,k] A[j,k]^T for tile columns j in [c0,c1), rows i >= j
  void update(int c0, int c1, int ka, int kb, cudaStream_t st) {
    if (c1 <= c0 || kb <= ka) return;
    GemmArgs a = base(kb - 1);
    a.j0 = c0; a.j1 = c1; a.i_off = 0;
    a.ka0 = ka * TILE; a.kb0 = ka * TILE; a.nk = (kb - ka) * TILE / GEMM_KB; a.b_row0 = 0;
    a.epi = 1;
    a.sym_lower = (factor && h->tri_skip) ? 1 : 0;
    launch(a, m.mapA, m.mapA, st);
  }
  void panel(int kb, int kend, cudaStream_t st) {
    for (int k = kb; k < kend; ++k) {
      if (factor) {
        TilePotrfArgs t{};
        t.A = m.A; t.lda = m.ld; t.a_batch_stride = m.batch_stride; t.k = k;
        t.Dinv = m.Dinv; t.d_batch_stride = m.dinv_bs;
        t.diag = m.diag; t.diag_batch_stride = m.diag_bs; t.info = m.info; t.pdl = pdl;
        if (m.rhs_r) {
          GPB_REQUIRE(tile_potrf_fuses_rhs(), "a right-hand side riding on the factorisation needs potrf_variant 3");
          t.rhs_r = m.rhs_r; t.rhs_z = m.rhs_z; t.rhs_bs = m.rhs_bs; t.rhs_gs = m.rhs_gs; t.rhs_zbs = m.rhs_zbs;
        }
        launch_tile_potrf_inv(t, m.batch, st);
        ++h->launches;
      }
      trsm(k, st);
      update(k + 1, kend, k, k + 1, st);
    }
  }
};

}  // namespace

void chol_trailing_update(gpb_handle* h, FactorMat& m, int c0, int c1, int ka, int kb) {
  const int nt = static_cast<int>(m.n_pad / TILE);
  const int R = static_cast<int>((m.rows_total + TILE - 1) / TILE);
  Sweep s{h, m, true, nt, R};
  s.update(c0, c1, ka, kb, h->s0);
}

void chol_sweep(gpb_handle* h, FactorMat& m, bool factor) {
  SweepPlan plan;
  plan.factor = factor;
  chol_sweep(h, m, plan);
}

void chol_sweep(gpb_handle* h, FactorMat& m, const SweepPlan& plan) {
  const bool factor = plan.factor;
  const int nt = static_cast<int>(m.n_pad / TILE);
  const int R = static_cast<int>((m.rows_total + TILE - 1) / TILE);
  Sweep s{h, m, factor, nt, R};
  s.x0 = plan.extra_tile0 >= 0 ? plan.extra_tile0 : nt;
  s.xn = plan.extra_tiles > 0 ? plan.extra_tiles : R - s.x0;
  s.grow = plan.grow;
  // outer block width: wide blocks amortise the per-tile turnover of the trailing update (K = 128 nb),
  // narrow ones keep the panel short where it cannot be hidden (measured: profiles/r01_tune_potrf.json).
  // With nb_tiles == 0 the width follows the REMAINING matrix: 4 tiles while the trailing update is long
  // enough to hide a 4-tile panel, then 2, then 1 in the panel-bound tail.
  // batches of small matrices (fewer than 24 tile columns) are throughput problems: plain order, blocks of 4 tiles
  // (round 1, 46 us diagonal tile: 2-tile blocks, 1024 x N=2048 124.7 vs 126.1 ms; round 2, 26 us tile: 4-tile blocks
  // 112.7 vs 114.0 ms; from N=4096 on the look-ahead schedule wins)
  const bool batch_plain = m.batch > 1 && (m.batch > h->la_max_batch || nt < 24);
  auto width_at = [&](int kb) {
    if (h->nb_tiles > 0) return h->nb_tiles;
    if (batch_plain) return h->batch_plain_width;
    // a small batch of big problems: the trailing update holds batch x the tiles, so the panel hides as it
    // would behind a single matrix sqrt(batch) times larger
    int rem = nt - kb;
    if (m.batch > 1) rem = static_cast<int>(rem * sqrt(static_cast<double>(m.batch)));
    if (h->nb_switch8 > 0 && rem >= h->nb_switch8) return 8;
    return rem >= h->nb_switch4 ? 4 : (rem >= h->nb_switch2 ? 2 : 1);
  };
  // without a symmetric part to factor there is no panel critical path: plain order
  // (up to 12 tile columns the second stream costs more than it hides: N = 1024 0.595 vs 0.622 ms)
  const bool la = h->lookahead && factor && !batch_plain && nt > width_at(0) && nt > 12;
  // Programmatic dependent launch pays where the chain of small kernels IS the run time (small matrices, thin
  // sweeps: +3 % at N <= 2048, thin appends); next to a look-ahead trailing update the early-resident waiters
  // take SM slots from it (measured: -4 % at N = 8192 / 16384), so it stays off there.
  s.pdl = !la || nt <= h->pdl_max_tiles;

  if (!la) {
    for (int kb = 0; kb < nt;) {
      const int nb = width_at(kb);
      const int kend = kb + nb < nt ? kb + nb : nt;
      s.panel(kb, kend, h->s0);
      s.update(kend, nt, kb, kend, h->s0);
      kb = kend;
    }
    return;
  }

  int kb = 0;
  int kend = width_at(0) < nt ? width_at(0) : nt;

  // ---- chunked schedule while the block is 4 tiles wide ------------------------------------------------
  // One launch per step serialises whole trailing updates: every launch ends in a partly filled wave, and step
  // s+1 cannot start a single tile before the last tile of step s is done.  Tiles of different column chunks are
  // independent, so here the update of step s is issued as chunks of 4 tile columns, chunk c always on stream
  // su[c % S]: the in-order stream gives "step s before step s+1" per chunk for free, one event per step makes
  // the panel visible to the update streams, one event hands the next panel's chunk to the panel stream.  CTAs of
  // several launches share the machine, so partial waves fill up and the panel chain runs ahead as far as its own
  // chunk allows.  Same kernels, same per-tile order of operations: results are bitwise identical.
  const int S = h->dag_streams;
  const bool dag = S > 0 && m.batch == 1 && h->nb_tiles == 0 && nt >= h->dag_min_tiles;
  s.panel(0, kend, h->s0);
  // one phase per block width (8, then 4): chunk -> stream is fixed inside a phase, phases are separated by a join
  while (dag && kend < nt && width_at(kend) >= h->dag_min_width) {
    const int CW = width_at(kend);
    while (static_cast<int>(h->su.size()) < S) {
      cudaStream_t st;
      GPB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
      h->su.push_back(st);
    }
    cudaEvent_t P = h->next_event();                       // panel [kb, kend) is done on s0
    GPB_CUDA(cudaEventRecord(P, h->s0));
    GPB_CUDA(cudaStreamWaitEvent(h->s1, P, 0));
    while (kend < nt && width_at(kend) == CW) {
      const int nend = kend + CW < nt ? kend + CW : nt;
      cudaEvent_t Pn = P;
      for (int i = 0; i < S; ++i) GPB_CUDA(cudaStreamWaitEvent(h->su[i], P, 0));
      for (int c = kend; c < nt; c += CW) {
        cudaStream_t st = h->su[(c / CW) % S];
        s.big_tiles = h->dag_big_tiles != 0;
        s.update(c, c + CW < nt ? c + CW : nt, kb, kend, st);
        s.big_tiles = false;
        if (c == kend) {                                   // the next panel's own chunk: hand it to the panel stream
          cudaEvent_t U = h->next_event();
          GPB_CUDA(cudaEventRecord(U, st));
          GPB_CUDA(cudaStreamWaitEvent(h->s1, U, 0));
          s.panel(kend, nend, h->s1);
          Pn = h->next_event();
          GPB_CUDA(cudaEventRecord(Pn, h->s1));
        }
      }
      P = Pn;
      kb = kend;
      kend = nend;
    }
    // join: everything issued so far becomes visible to the handle's stream
    GPB_CUDA(cudaStreamWaitEvent(h->s0, P, 0));
    for (int i = 0; i < S; ++i) {
      cudaEvent_t e = h->next_event();
      GPB_CUDA(cudaEventRecord(e, h->su[i]));
      GPB_CUDA(cudaStreamWaitEvent(h->s0, e, 0));
    }
  }
  // The panel stream carries the whole critical chain - update of the next panel's columns, diagonal tile, TRSM -
  // as consecutive launches of ONE stream (no event hop inside the chain, programmatic dependent launch applies);
  // the main stream only applies each finished panel to the columns further right.  Per step two events: `er`
  // (main -> panel: the previous rest-update has reached the next panel's columns) and `ep` (panel -> main).
  if (h->chain_on_panel_stream) {
    cudaEvent_t er = h->next_event();
    GPB_CUDA(cudaEventRecord(er, h->s0));              // panel [kb, kend) and everything before it
    while (kend < nt) {
      const int nbn = width_at(kend);
      const int nend = kend + nbn < nt ? kend + nbn : nt;
      // the tail of a big matrix is a small matrix: from here on the chain of small launches is the run time again
      if (h->pdl_tail && nt - kend <= h->pdl_max_tiles) s.pdl = true;
      GPB_CUDA(cudaStreamWaitEvent(h->s1, er, 0));
      s.update(kend, nend, kb, kend, h->s1);           // columns of the next panel ...
      s.panel(kend, nend, h->s1);                      // ... and the panel itself, back to back
      cudaEvent_t ep = h->next_event();
      GPB_CUDA(cudaEventRecord(ep, h->s1));
      s.update(nend, nt, kb, kend, h->s0);             // meanwhile the rest of the update
      er = h->next_event();
      GPB_CUDA(cudaEventRecord(er, h->s0));
      GPB_CUDA(cudaStreamWaitEvent(h->s0, ep, 0));
      kb = kend;
      kend = nend;
    }
    return;
  }
  while (kend < nt) {
    const int nbn = width_at(kend);
    const int nend = kend + nbn < nt ? kend + nbn : nt;
    s.update(kend, nend, kb, kend, h->s0);            // columns of the next panel first
    cudaEvent_t eu = h->next_event();
    GPB_CUDA(cudaEventRecord(eu, h->s0));
    GPB_CUDA(cudaStreamWaitEvent(h->s1, eu, 0));
    s.panel(kend, nend, h->s1);                        // next panel on the side stream ...
    cudaEvent_t ep = h->next_event();
    GPB_CUDA(cudaEventRecord(ep, h->s1));
    s.update(nend, nt, kb, kend, h->s0);               // ... while the rest of the update runs
    GPB_CUDA(cudaStreamWaitEvent(h->s0, ep, 0));
    kb = kend;
    kend = nend;
  }
}

}  // namespace gpb
