// dmma_gemm.cu - the kernel the factorisation spends its time in.
//
// C(it,jt) (op)= sum_k A(it,k) * B(jt,k)^T on tiles of row-major fp64 matrices, used for
//   - the trailing SYRK/GEMM update of the blocked Cholesky (np.linalg.cholesky, GPr.py:62),
//   - the panel TRSM written as a product with the inverted diagonal tile,
//   - the fused forward solves (rows appended under K), TRTRI/LAUUM of the gradient stage.
//
// B200 mapping
//   * tcgen05 has no f64 kind: the FP64 tensor path on sm_100a is the warp-level
//     mma.sync.m8n8k4 (SASS DMMA.8x8x4), operands in registers.
//   * operand slabs (tile rows x 16 doubles = 128-byte rows) are brought in by TMA
//     (cp.async.bulk.tensor, SASS UTMALDG) with the 128B swizzle through a 4-stage
//     full/empty mbarrier ring.  The producer duty belongs to the elected lane of warp 0 (registers
//     are allocated per four warps: a ninth warp would cap the DMMA warps at 168 registers).
//   * work list: (batch entry, tile) pairs, walked with a grid stride; the slab ring runs ACROSS tile
//     boundaries, so a CTA that owns several work items prefetches the next one during the current
//     epilogue.  By default the grid has one CTA per work item (the hardware scheduler balances SMs
//     and lets the look-ahead's high-priority panel kernels in between CTAs); a persistent grid is an
//     option (measured: +3 % at K=512, -5 % at K=8192, look-ahead starved).
//   * fragment rows are taken with a stride of two tile rows ("parity" fragments): under the
//     128B swizzle the 16 lanes of a half warp then read 8 distinct 16-byte chunks over 4 rows
//     = all 32 banks once, so every 64-bit fragment load is conflict free.
//   * instantiations: 128x64 (4 warps of 64x32, two CTAs per SM, second one phase-shifted) for the big
//     trailing updates - per 16-wide slab a warp issues 128 DMMA for 48 LDS.64, the FP64 tensor pipe
//     is the only busy unit (92 % active at K=512); 128x128 (8 warps, 1 CTA/SM); 64x64 (4 warps of
//     32x32, 3 CTAs/SM) for launches on the panel's critical path or too small to fill 148 SMs, and for every
//     short-k update of a batch of small matrices; 64x128 (8 warps, 2 CTAs/SM) for the in-place panel TRSM (a CTA
//     must own all 128 output columns of its rows); 32x128 and 32x64 (8 warps) for the TRSM and the next-column
//     update of a small matrix, where the launch is one wave and its duration is the latency of ONE CTA.
//   * round 2: a triangular B operand (the inverted diagonal tile) lets a warp stop after the k slabs its columns
//     need (b_tri), symmetric updates skip warp tiles above the diagonal (sym_lower), and the panel-TRSM shapes have
//     a second instantiation (GEMV) whose epilogue takes the just-computed columns out of the running right-hand side
//     of the forward substitution.
#include "gpb_kernels.cuh"

namespace gpb {

__device__ __forceinline__ void decode_tile(const GemmArgs& p, int idx, int& it, int& jt) {
  if (p.tri) {
    // column c (jt = j0 + c) holds H - c tiles, H = R - j0 - i_off; off(c) = c*H - c(c-1)/2
    const int H = p.R - p.j0 - p.i_off;
    const double b = 2.0 * H + 1.0;
    int c = static_cast<int>((b - sqrt(b * b - 8.0 * idx)) * 0.5);
    if (c < 0) c = 0;
    while ((c + 1) * H - (c + 1) * c / 2 <= idx) ++c;
    while (c * H - c * (c - 1) / 2 > idx) --c;
    jt = p.j0 + c;
    it = jt + p.i_off + (idx - (c * H - c * (c - 1) / 2));
  } else {
    const int nrows = p.R - p.i0;
    jt = p.j0 + idx / nrows;
    it = p.i0 + idx % nrows;
  }
}

// BM x BN output tile per CTA-tile.  The region is described in BM x BM tiles (square BN >= BM
// configurations: BM x BN); when BN < BM a region tile is cut into BM/BN column slices.
// GEMV: the instantiation carries the fused substitution step (GemmArgs::gemv_*); only the panel-TRSM shapes (BN = 128)
// have one, so the update kernels - where the time is - stay exactly as they were.
template <int BM, int BN, int WM, int WN, int MINB, bool GEMV = false>
__global__ void __launch_bounds__(WM * WN * 32, MINB)
dmma_gemm_nt_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                    const GemmArgs p) {
  constexpr int NCW = WM * WN;               // warps (all of them DMMA consumers)
  constexpr int BMW = BM / WM;               // warp tile rows
  constexpr int BNW = BN / WN;               // warp tile columns
  constexpr int GM = BMW / 16;               // 16-row groups per warp (two parity fragments each)
  constexpr int GN = BNW / 16;
  constexpr int SLAB_A = BM * GEMM_KB * 8;   // bytes of one operand slab
  constexpr int SLAB_B = BN * GEMM_KB * 8;
  constexpr int STAGE = SLAB_A + SLAB_B;
  constexpr int NSPLIT = (BN < BM) ? BM / BN : 1;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * STAGE);
  uint64_t* empty = full + GEMM_STAGES;

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int per_batch = p.ntiles * NSPLIT;   // CTA-tiles per batch entry
  const int total = per_batch * p.nbatch;    // work items of this launch: (batch entry, CTA-tile)

  // CTAs that share an SM start together and then run in lockstep - both in their prologue, both in their
  // epilogue at the same time, so nothing overlaps.  The second resident CTA of every SM (first wave only)
  // waits half a tile time; the pair then stays out of phase for the rest of the launch.
  if (MINB >= 2 && p.stagger_clk > 0 && blockIdx.x >= static_cast<unsigned>(p.num_sms) &&
      blockIdx.x < 2u * static_cast<unsigned>(p.num_sms)) {
    const long long t0 = clock64();
    while (clock64() - t0 < p.stagger_clk) __nanosleep(2000);
  }

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCW);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();                                // operands and C come from the preceding kernels of the stream

  // ---------------- producer state (only thread 0 uses it) ----------------
  const bool producer = (threadIdx.x == 0);
  int iss_tile = blockIdx.x;                 // CTA-tile the next issued slab belongs to
  int iss_s = 0, iss_nk = 0, iss_ka = 0, iss_kb = 0, iss_arow = 0, iss_brow = 0, iss_batch = 0;
  uint32_t n_issued = 0;                     // slabs issued so far (ring position)
  auto open_issue_tile = [&]() {
    if (iss_tile < total) {
      int ti, tj;
      iss_batch = iss_tile / per_batch;
      const int w = iss_tile - iss_batch * per_batch;
      decode_tile(p, w / NSPLIT, ti, tj);
      tj = tj * NSPLIT + w % NSPLIT;
      iss_nk = p.k_from_row ? (p.k_end - ti * BM) / GEMM_KB : p.nk;
      iss_ka = p.k_from_row ? ti * BM : p.ka0;
      iss_kb = p.k_from_row ? ti * BM : p.kb0;
      iss_arow = p.a_row0 + ti * BM;
      iss_brow = p.b_row0 + tj * BN;
      iss_s = 0;
    }
  };
  auto issue_next = [&]() {                  // caller guarantees iss_tile < total and a free slot
    const int st = n_issued % GEMM_STAGES;
    uint8_t* dst = smem + st * STAGE;
    mbar_arrive_expect_tx(&full[st], STAGE);
    tma_load_3d(dst, &mapA, &full[st], iss_ka + GEMM_KB * iss_s, iss_arow, iss_batch);
    tma_load_3d(dst + SLAB_A, &mapB, &full[st], iss_kb + GEMM_KB * iss_s, iss_brow, iss_batch);
    ++n_issued;
    if (++iss_s == iss_nk) {
      iss_tile += gridDim.x;
      open_issue_tile();
    }
  };
  if (producer) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    open_issue_tile();
    for (int s = 0; s < GEMM_STAGES && iss_tile < total; ++s) issue_next();
  }

  // ---------------- DMMA consumers ----------------
  // Warp w runs on scheduler w % 4.  With a triangular B tile the work of a warp grows with its column group wn, so
  // the second row of warps takes the column groups in reverse order: every scheduler then carries a light and a
  // heavy warp (WN = 4), instead of one scheduler carrying both warps that need all k slabs.  Only then: in the
  // plain update kernels the reversed order measured 1.3 % slower at N = 16384 (profiles/r02_ab_kernel_variants.txt).
  const int wm = warp / WN;
  const int wn = (p.b_tri && (wm & 1)) ? WN - 1 - warp % WN : warp % WN;
  const int g = lane >> 2;         // fragment row (A) / column (B)
  const int t = lane & 3;          // fragment k index
  const int th = t >> 1;

  // byte offsets inside a slab (row r, column c): r*128 + (((c>>1) ^ (r&7)) << 4) + (c&1)*8
  uint32_t a_off[2], b_off[2], xr[2];
#pragma unroll
  for (int par = 0; par < 2; ++par) {
    xr[par] = ((g & 3) << 1) | par;
    a_off[par] = (wm * BMW + 2 * g + par) * 128 + (t & 1) * 8;
    b_off[par] = (wn * BNW + 2 * g + par) * 128 + (t & 1) * 8;
  }

  uint32_t n_done = 0;             // slabs consumed so far by this warp (ring position)
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int it, jt;
    const int batch = tile / per_batch;
    {
      const int w = tile - batch * per_batch;
      decode_tile(p, w / NSPLIT, it, jt);
      jt = jt * NSPLIT + w % NSPLIT;
    }
    const int nk = p.k_from_row ? (p.k_end - it * BM) / GEMM_KB : p.nk;
    // slabs this warp multiplies (b_tri) and whether its tile is needed at all (sym_lower)
    const int nk_warp = p.b_tri ? min(nk, ((wn + 1) * BNW + GEMM_KB - 1) / GEMM_KB) : nk;
    const bool dead = p.sym_lower && (it * BM + wm * BMW + BMW - 1 < jt * BN + wn * BNW);

    if (p.epi == 1 && !dead) {
      // pull this warp's piece of C towards L2 while the k loop runs (128-byte lines)
      const double* Cp = p.C + static_cast<int64_t>(batch) * p.c_batch_stride;
      const int64_t rb = static_cast<int64_t>(it) * BM + wm * BMW;
      const int64_t cb = static_cast<int64_t>(jt) * BN + wn * BNW;
      constexpr int LPR = BNW / 16;                       // lines per row
#pragma unroll
      for (int q = 0; q < BMW * LPR / 32; ++q) {
        const int idx = lane + 32 * q;
        const int64_t row = rb + idx / LPR;
        if (row < p.rows_total) prefetch_l2(Cp + row * p.ldc + cb + (idx % LPR) * 16);
      }
    }

    double acc[2 * GM][2 * GN][2];
#pragma unroll
    for (int i = 0; i < 2 * GM; ++i)
#pragma unroll
      for (int j = 0; j < 2 * GN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // The math loop is kept free of the skip logic (a predicate around the DMMA block cost 3 % at N = 16384): a warp
    // multiplies its first n_math slabs, then only keeps the slab ring turning for the rest (it must stay in step with
    // the ring: wait for the slab, release it; warp 0 also keeps refilling).
    const int n_math = dead ? 0 : nk_warp;
    for (int s = 0; s < n_math; ++s) {
      if (producer && n_done >= 1 && iss_tile < total) {
        // refill the slot of the slab this warp has just left, once every warp has released it;
        // the slab issued here may already belong to the next tile
        mbar_wait(&empty[(n_done - 1) % GEMM_STAGES], ((n_done - 1) / GEMM_STAGES) & 1);
        issue_next();
      }
      __syncwarp();
      const int st = n_done % GEMM_STAGES;
      mbar_wait(&full[st], (n_done / GEMM_STAGES) & 1);
      const uint8_t* sa = smem + st * STAGE;
      const uint8_t* sb = sa + SLAB_A;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        double af[2 * GM], bf[2 * GN];
#pragma unroll
        for (int par = 0; par < 2; ++par) {
          const uint32_t chunk = ((2 * kk + th) ^ xr[par]) << 4;
#pragma unroll
          for (int grp = 0; grp < GM; ++grp)
            af[grp * 2 + par] = *reinterpret_cast<const double*>(sa + a_off[par] + grp * 2048 + chunk);
#pragma unroll
          for (int grp = 0; grp < GN; ++grp)
            bf[grp * 2 + par] = *reinterpret_cast<const double*>(sb + b_off[par] + grp * 2048 + chunk);
        }
#pragma unroll
        for (int mi = 0; mi < 2 * GM; ++mi)
#pragma unroll
          for (int nj = 0; nj < 2 * GN; ++nj) dmma884(acc[mi][nj][0], acc[mi][nj][1], af[mi], bf[nj]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      ++n_done;
    }
    for (int s = n_math; s < nk; ++s) {
      if (producer && n_done >= 1 && iss_tile < total) {
        mbar_wait(&empty[(n_done - 1) % GEMM_STAGES], ((n_done - 1) / GEMM_STAGES) & 1);
        issue_next();
      }
      __syncwarp();
      const int st = n_done % GEMM_STAGES;
      mbar_wait(&full[st], (n_done / GEMM_STAGES) & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
      ++n_done;
    }

    if (GEMV && p.gemv_r != nullptr) {
      // r_g[row] -= X[row][16 g .. 16 g + 15] . z[16 g ..] for the rows of this CTA-tile (BN = 128: the tile spans the
      // whole column block).  Per 16-column group the lane's four products in a fixed order, then the four lanes of a
      // row by shuffle; lane t == 0 owns (group, row).  No reduction across warps: the sum must not depend on the warp
      // layout of the instantiation (tests/_sweep_nccl_worker.py), and a first version that reduced the groups through
      // shared memory behind CTA barriers gave irreproducible results under the look-ahead schedules at N >= 9216
      // (cause not found; this form is checked by tools/fused_rhs_determinism.py and tests/test_gpu_config_sizes.py).
      const double* zt = p.gemv_z + static_cast<int64_t>(batch) * p.gemv_zbs + wn * BNW;
      double* rb = p.gemv_r + static_cast<int64_t>(batch) * p.gemv_bs;
#pragma unroll
      for (int gn = 0; gn < GN; ++gn) {
        const double z00 = zt[16 * gn + 4 * t], z10 = zt[16 * gn + 4 * t + 1];          // (pn, e) = (0,0), (1,0)
        const double z01 = zt[16 * gn + 4 * t + 2], z11 = zt[16 * gn + 4 * t + 3];      // (0,1), (1,1)
        double* rg = rb + static_cast<int64_t>(wn * GN + gn) * p.gemv_gs;
#pragma unroll
        for (int mi = 0; mi < 2 * GM; ++mi) {
          double sdot = acc[mi][gn * 2][0] * z00;
          sdot = fma(acc[mi][gn * 2 + 1][0], z10, sdot);
          sdot = fma(acc[mi][gn * 2][1], z01, sdot);
          sdot = fma(acc[mi][gn * 2 + 1][1], z11, sdot);
          sdot += __shfl_xor_sync(0xffffffffu, sdot, 1);
          sdot += __shfl_xor_sync(0xffffffffu, sdot, 2);
          const int64_t row = static_cast<int64_t>(it) * BM + wm * BMW + 16 * (mi >> 1) + 2 * g + (mi & 1);
          if (t == 0 && row < p.rows_total) rg[row] -= sdot;
        }
      }
    }

    // ---------------- epilogue: each lane owns 4 consecutive columns per (row, column group) -------
    // fragment (grp_m, par_m) row g  -> tile row 16 grp_m + 2 g + par_m
    // fragment (grp_n, par_n) col 2t+e -> tile col 16 grp_n + 4 t + 2 e + par_n
    double* Cb = p.C + static_cast<int64_t>(batch) * p.c_batch_stride;
    const int64_t row_base = static_cast<int64_t>(it) * BM + wm * BMW;
    const int64_t col_base = static_cast<int64_t>(jt) * BN + wn * BNW;
#pragma unroll
    for (int gm = 0; gm < GM; ++gm) {
#pragma unroll
      for (int pm = 0; pm < 2; ++pm) {
        const int64_t row = row_base + 16 * gm + 2 * g + pm;
        if (row < p.rows_total && !dead && (p.epi != 2 || acc[0][0][0] == 1.2345e300)) {   // epi 2: measurement only, no stores
          const int mi = gm * 2 + pm;
#pragma unroll
          for (int gn = 0; gn < GN; ++gn) {
            double* ptr = Cb + row * p.ldc + col_base + 16 * gn + 4 * t;
            double2 lo = make_double2(acc[mi][gn * 2][0], acc[mi][gn * 2 + 1][0]);
            double2 hi = make_double2(acc[mi][gn * 2][1], acc[mi][gn * 2 + 1][1]);
            if (p.epi == 1) {
              const double2 c0 = *reinterpret_cast<const double2*>(ptr);
              const double2 c1 = *reinterpret_cast<const double2*>(ptr + 2);
              lo.x = c0.x - lo.x; lo.y = c0.y - lo.y;
              hi.x = c1.x - hi.x; hi.y = c1.y - hi.y;
            }
            *reinterpret_cast<double2*>(ptr) = lo;
            *reinterpret_cast<double2*>(ptr + 2) = hi;
          }
        }
      }
    }
  }
}

int gemm_region_tiles(const GemmArgs& a) {
  const int ncols = a.j1 - a.j0;
  if (ncols <= 0) return 0;
  if (a.tri) {
    const int H = a.R - a.j0 - a.i_off;
    if (H - (ncols - 1) <= 0) return -1;
    return ncols * H - ncols * (ncols - 1) / 2;
  }
  if (a.R - a.i0 <= 0) return 0;
  return ncols * (a.R - a.i0);
}

constexpr int smem_bytes(int bm, int bn) { return GEMM_STAGES * (bm + bn) * GEMM_KB * 8 + 1024 + 256; }

static int g_num_sms = 148;
static int g_persistent_waves = 0;
void dmma_gemm_set_persistent(int waves) { g_persistent_waves = waves < 0 ? 0 : waves; }
int g_pdl = 1;
int g_capturing = 0;
void dmma_gemm_set_pdl(int mode) { g_pdl = mode < 0 ? 0 : mode; }
static int g_fine_warps = 1;
void dmma_gemm_set_fine_warps(int on) { g_fine_warps = on != 0; }
static int g_stagger = 1;
void dmma_gemm_set_stagger(int on) { g_stagger = on != 0; }
static int g_trsm_persist = 2;   // (1024 x N=2048: 103.23 -> 102.87 ms, 128 problems 13.39 -> 13.30; same bits)
void dmma_gemm_set_trsm_persist(int waves) { g_trsm_persist = waves < 0 ? 0 : waves; }
static int g_trsm_balance = 1;
void dmma_gemm_set_trsm_balance(int mode) { g_trsm_balance = mode; }

void dmma_gemm_init() {
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<128, 128, 2, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(128, 128)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<64, 64, 2, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(64, 64)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<64, 128, 2, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(64, 128)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<64, 64, 2, 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(64, 64)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<64, 128, 2, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(64, 128)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<128, 64, 2, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(128, 64)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<32, 128, 2, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(32, 128)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<128, 128, 2, 4, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(128, 128)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<64, 128, 2, 4, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(64, 128)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<64, 128, 2, 2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(64, 128)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<32, 128, 2, 4, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(32, 128)));
  GPB_CUDA(cudaFuncSetAttribute(dmma_gemm_nt_kernel<32, 64, 2, 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                smem_bytes(32, 64)));
  int dev = 0, sms = 0;
  GPB_CUDA(cudaGetDevice(&dev));
  GPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sms > 0) g_num_sms = sms;
}

void launch_dmma_gemm(const CUtensorMap& mapA, const CUtensorMap& mapB, GemmArgs a, int batch,
                      cudaStream_t st, int tile) {
  const int ntiles = gemm_region_tiles(a);
  GPB_REQUIRE(ntiles >= 0, "dmma_gemm: empty column in trapezoid region");
  if (ntiles == 0 || (a.nk == 0 && !a.k_from_row)) return;
  a.ntiles = ntiles;
  // Grid: by default one CTA per work item - the hardware scheduler balances SMs of unequal speed and
  // lets the high-priority panel kernels of the look-ahead in between CTAs.  g_persistent_waves > 0
  // caps the grid at that many resident waves instead (CTAs then walk the list with a grid stride
  // and prefetch across tile boundaries; measured: +3 % at K=512, -5 % at K=8192, look-ahead starved).
  a.nbatch = batch < 1 ? 1 : batch;
  a.num_sms = g_num_sms;
  {
    // half of the time a CTA-tile takes when two CTAs share the FP64 tensor pipe: nk slabs x 2048 clocks
    const long long nk = a.k_from_row ? (a.k_end / GEMM_KB) / 2 : a.nk;
    a.stagger_clk = (g_stagger && !a.no_stagger && static_cast<int64_t>(ntiles) * a.nbatch >= 2LL * g_num_sms) ? nk * 2048 : 0;
  }
  auto grid_for = [&](int cta_tiles, int per_sm) {
    const int64_t work = static_cast<int64_t>(cta_tiles) * a.nbatch;
    int64_t gx = work;
    const int waves = (a.b_tri && g_trsm_persist > 0) ? g_trsm_persist : g_persistent_waves;
    if (waves > 0) {
      const int64_t resident = static_cast<int64_t>(g_num_sms) * per_sm * waves;
      if (resident < gx) gx = resident;
    }
    return dim3(static_cast<unsigned>(gx), 1, 1);
  };
  // programmatic dependent launch for the launches that fit the machine at once (the dependent chains of the panel
  // and of the sweeps); a multi-wave trailing update gains nothing and its early-resident successor would sit in
  // slots the look-ahead panel wants
  const bool small = static_cast<int64_t>(ntiles) * a.nbatch <= 2LL * g_num_sms;
  const bool pdl = g_pdl == 2 || (g_pdl == 1 && small && a.pdl != 0);
  const bool gemv = a.gemv_r != nullptr;
  GPB_REQUIRE(!gemv || tile == 128 || tile == 64128 || tile == 32128, "dmma_gemm: the fused substitution step needs a 128-column tile");
  if (tile == 128 && gemv) {
    launch_chain(dmma_gemm_nt_kernel<128, 128, 2, 4, 1, true>, grid_for(ntiles, 1), dim3(8 * 32), smem_bytes(128, 128), st, pdl,
                 mapA, mapB, a);
  } else if (tile == 128) {
    launch_chain(dmma_gemm_nt_kernel<128, 128, 2, 4, 1>, grid_for(ntiles, 1), dim3(8 * 32), smem_bytes(128, 128), st, pdl,
                 mapA, mapB, a);
  } else if (tile == 12864) {
    // region in 128-tiles, each cut into two 128 x 64 CTA-tiles; two CTAs share an SM
    // (mapA with 128-row boxes, mapB with 64-row boxes)
    launch_chain(dmma_gemm_nt_kernel<128, 64, 2, 2, 2>, grid_for(2 * ntiles, 2), dim3(4 * 32), smem_bytes(128, 64), st, pdl,
                 mapA, mapB, a);
  } else if (tile == 32128) {
    // 32 x 128, 8 warps of 16 x 32: rows in units of 32, columns in units of 128 (the in-place panel TRSM of a small matrix)
    if (gemv)
      launch_chain(dmma_gemm_nt_kernel<32, 128, 2, 4, 2, true>, grid_for(ntiles, 2), dim3(8 * 32), smem_bytes(32, 128), st, pdl,
                   mapA, mapB, a);
    else
      launch_chain(dmma_gemm_nt_kernel<32, 128, 2, 4, 2>, grid_for(ntiles, 2), dim3(8 * 32), smem_bytes(32, 128), st, pdl,
                   mapA, mapB, a);
  } else if (tile == 3264) {
    // 32 x 64, 8 warps of 16 x 16: rows in units of 32, columns in units of 64 (single-column update on the critical chain)
    launch_chain(dmma_gemm_nt_kernel<32, 64, 2, 4, 3>, grid_for(ntiles, 3), dim3(8 * 32), smem_bytes(32, 64), st, pdl,
                 mapA, mapB, a);
  } else if (tile == 64) {
    // A launch that leaves SMs with a single CTA is latency bound (one warp per scheduler: 1.34 us per 16-deep k
    // step for a lone 4-warp CTA, 40 % of the pipe rate): such launches - the panel's critical chain - use 8 warps
    // with half-size warp tiles, two warps per scheduler covering each other's fragment loads.
    if (g_fine_warps && static_cast<int64_t>(ntiles) * a.nbatch <= g_num_sms)
      launch_chain(dmma_gemm_nt_kernel<64, 64, 2, 4, 3>, grid_for(ntiles, 3), dim3(8 * 32), smem_bytes(64, 64), st, pdl,
                   mapA, mapB, a);
    else
      launch_chain(dmma_gemm_nt_kernel<64, 64, 2, 2, 3>, grid_for(ntiles, 3), dim3(4 * 32), smem_bytes(64, 64), st, pdl,
                   mapA, mapB, a);
  } else {
    // 64 x 128: rows in units of 64, columns in units of 128 (mapA with 64-row boxes, mapB with 128-row boxes)
    // triangular B (panel TRSM): the light and the heavy column groups have to be spread over the four schedulers.
    // 8 warps (2 x 4, second row mirrored) do that inside one CTA; with 4 warps (2 x 2) a scheduler holds one warp of
    // each resident CTA and both are of the same kind.  128-problem slice of the sweep: 14.32 -> 13.76 ms, same bits
    // (tools/sweep_balance.py; mirroring the assignment in every other 4-warp CTA instead: no gain).
    const bool fine = (g_fine_warps && static_cast<int64_t>(ntiles) * a.nbatch <= g_num_sms) || (a.b_tri && g_trsm_balance);
    if (fine && gemv)
      launch_chain(dmma_gemm_nt_kernel<64, 128, 2, 4, 2, true>, grid_for(ntiles, 2), dim3(8 * 32), smem_bytes(64, 128), st, pdl,
                   mapA, mapB, a);
    else if (fine)
      launch_chain(dmma_gemm_nt_kernel<64, 128, 2, 4, 2>, grid_for(ntiles, 2), dim3(8 * 32), smem_bytes(64, 128), st, pdl,
                   mapA, mapB, a);
    else if (gemv)
      launch_chain(dmma_gemm_nt_kernel<64, 128, 2, 2, 2, true>, grid_for(ntiles, 2), dim3(4 * 32), smem_bytes(64, 128), st, pdl,
                   mapA, mapB, a);
    else
      launch_chain(dmma_gemm_nt_kernel<64, 128, 2, 2, 2>, grid_for(ntiles, 2), dim3(4 * 32), smem_bytes(64, 128), st, pdl,
                   mapA, mapB, a);
  }
}

// ---------------------------------------------------------------------------------------
// Pipe-rate micro-benchmarks: fix the FP64 roofline denominator on the box itself.
// ---------------------------------------------------------------------------------------
constexpr int MB_ITERS = 4096;
__global__ void __launch_bounds__(256) mb_dmma_kernel(double* sink) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < MB_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) sink[0] = s;
}
__global__ void __launch_bounds__(256) mb_dfma_kernel(double* sink) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < MB_ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) sink[0] = s;
}
double launch_microbench(int kind, double* sink, cudaStream_t st) {
  const int blocks = 148 * 8, threads = 256;
  if (kind == 0) {
    mb_dmma_kernel<<<blocks, threads, 0, st>>>(sink);
    GPB_CUDA(cudaGetLastError());
    return 2.0 * 256.0 * 16.0 * MB_ITERS * (threads / 32) * static_cast<double>(blocks);
  }
  mb_dfma_kernel<<<blocks, threads, 0, st>>>(sink);
  GPB_CUDA(cudaGetLastError());
  return 2.0 * 16.0 * MB_ITERS * threads * static_cast<double>(blocks);
}

}  // namespace gpb
