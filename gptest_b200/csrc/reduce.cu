// reduce.cu - finishing reductions of the regression path (deterministic fixed-order trees).
#include "gpb_kernels.cuh"

namespace gpb {

// fixed-order block reduction: every thread's partial goes through the same tree every run
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
#pragma unroll
  for (int s = NT / 2; s > 0; s >>= 1) {
    if (t < s) sh[t] += sh[t + s];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

// GPr.py:65-68: err_y = 0.5 y' K^-1 y = 0.5 |L^-1 y|^2 ; det = sum(log(diag(L))) ; n log(2 pi)/2
__global__ void __launch_bounds__(512) nlml_finish_kernel(const double* __restrict__ z, int64_t z_bs,
                                                          const double* __restrict__ diag, int64_t d_bs,
                                                          int64_t n_pad, int64_t n_valid,
                                                          double* __restrict__ out) {
  __shared__ double sh[512];
  const int b = blockIdx.x;
  const double* zb = z + b * z_bs;
  const double* db = diag + b * d_bs;
  double q = 0.0, l = 0.0;
  for (int64_t i = threadIdx.x; i < n_pad; i += 512) {
    const double zi = zb[i];
    q = fma(zi, zi, q);
    l += log(db[i]);
  }
  const double qs = block_sum<512>(q, sh);
  const double ls = block_sum<512>(l, sh);
  if (threadIdx.x == 0) out[b] = 0.5 * qs + ls + 0.5 * static_cast<double>(n_valid) * 1.8378770664093453;
}

void launch_nlml_finish(const double* z, int64_t z_bs, const double* diag, int64_t d_bs, int64_t n_pad,
                        int64_t n_valid, double* out, int batch, cudaStream_t st) {
  nlml_finish_kernel<<<batch, 512, 0, st>>>(z, z_bs, diag, d_bs, n_pad, n_valid, out);
  GPB_CUDA(cudaGetLastError());
}

// GPr.py:50-53 with V^T = Kzx L^-T and z = L^-1 y:  fz = V^T z,  cov = sf2 - rowsum(V^T * V^T)
__global__ void __launch_bounds__(256) predict_finish_kernel(const double* __restrict__ VT, int64_t ld,
                                                             const double* __restrict__ z, int64_t n_pad,
                                                             const double* __restrict__ hyp,
                                                             double* __restrict__ mean, double* __restrict__ var) {
  __shared__ double sh[256];
  const int64_t i = blockIdx.x;
  const double* row = VT + i * ld;
  double m = 0.0, s = 0.0;
  for (int64_t j = 2 * threadIdx.x; j < n_pad; j += 512) {
    const double2 v = *reinterpret_cast<const double2*>(row + j);
    const double2 zz = *reinterpret_cast<const double2*>(z + j);
    m = fma(v.x, zz.x, m);
    m = fma(v.y, zz.y, m);
    s = fma(v.x, v.x, s);
    s = fma(v.y, v.y, s);
  }
  const double ms = block_sum<256>(m, sh);
  const double ss = block_sum<256>(s, sh);
  if (threadIdx.x == 0) {
    mean[i] = ms;
    var[i] = hyp[0] - ss;
  }
}

void launch_predict_finish(const double* VT, int64_t ld, const double* z, int64_t n_pad, int64_t m,
                           const double* hyp_dev, double* mean, double* var, cudaStream_t st) {
  if (m == 0) return;
  predict_finish_kernel<<<static_cast<unsigned>(m), 256, 0, st>>>(VT, ld, z, n_pad, hyp_dev, mean, var);
  GPB_CUDA(cudaGetLastError());
}

// row `row_off / ld` of every batch entry <- y - mean (zero padded): the right-hand side stored as a row
__global__ void set_y_rows_kernel(double* A, int64_t batch_stride, int64_t row_off, const double* y,
                                  int64_t n, int64_t n_pad, double mean) {
  double* dst = A + blockIdx.y * batch_stride + row_off;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_pad;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = (i < n) ? y[i] - mean : 0.0;
}
void launch_set_y_rows(double* A, int64_t batch_stride, int64_t row_off, const double* y, int64_t n,
                       int64_t n_pad, double mean, int batch, cudaStream_t st) {
  dim3 grid(static_cast<unsigned>((n_pad + 255) / 256), batch);
  set_y_rows_kernel<<<grid, 256, 0, st>>>(A, batch_stride, row_off, y, n, n_pad, mean);
  GPB_CUDA(cudaGetLastError());
}

// One step of the blocked forward substitution  L x = r  with the inverted diagonal tiles (batched):
//   x_k = W_k r_k ;  r_i -= L[i][tile k columns] . x_k  for every row i below tile k.
// Replaces the appended y-row of the sweep where a whole 128-row tile for one right-hand side is too
// expensive (N = 2048 batched: 16 % of the GEMM work).  Reads L once: HBM bound.
__global__ void __launch_bounds__(256) trsv_l_step_kernel(const double* __restrict__ L, int64_t ld, int64_t l_bs,
                                                          const double* __restrict__ Dinv, int64_t d_bs, int k,
                                                          int64_t n_pad, double* __restrict__ r, int64_t r_bs,
                                                          double* __restrict__ x, int64_t x_bs, int groups) {
  __shared__ __align__(32) double rk[TILE];
  __shared__ __align__(32) double xk[TILE];
  const int t = threadIdx.x, b = blockIdx.y;
  pdl_trigger();
  pdl_wait();                                    // r comes from the previous step
  double* rb = r + b * r_bs;
  if (t < TILE) rk[t] = rb[k * TILE + t];
  __syncthreads();
  const int warp = t >> 5, lane = t & 31;
  {
    // x_k[row] = sum_{c <= row} W[row][c] r_k[c]: warp w takes rows 16w..16w+15, a row is one coalesced 1 KB read
    // (4 consecutive doubles per lane; lanes right of the diagonal skip the load: W is lower triangular), the 16
    // row loads of a warp are independent, fixed-order butterfly reduce
    const double* W = Dinv + b * d_bs + static_cast<int64_t>(k) * TILE * TILE;
    const double4 rv = *reinterpret_cast<const double4*>(&rk[4 * lane]);
    double4 wv[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int row = 16 * warp + q;
      wv[q] = (4 * lane <= row) ? *reinterpret_cast<const double4*>(W + row * TILE + 4 * lane) : make_double4(0.0, 0.0, 0.0, 0.0);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      double s = fma(wv[q].x, rv.x, fma(wv[q].y, rv.y, fma(wv[q].z, rv.z, wv[q].w * rv.w)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) xk[16 * warp + q] = s;
    }
  }
  __syncthreads();
  if (t < TILE && blockIdx.x == 0) x[b * x_bs + k * TILE + t] = xk[t];   // not in place: other CTAs of this launch still read r_k
  // rows below the tile: one warp per row, 4 consecutive doubles per lane (1 KB coalesced), fixed-order reduce
  const double4 xv = *reinterpret_cast<const double4*>(&xk[4 * lane]);
  const double* Lb = L + b * l_bs;
  // a CTA takes `groups` blocks of 64 rows: every CTA reads the 128 KB W tile first, so few long CTAs beat many short
  for (int gi = 0; gi < groups; ++gi)
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int64_t row = static_cast<int64_t>(k + 1) * TILE + (static_cast<int64_t>(blockIdx.x) * groups + gi) * 64 + warp * 8 + q;
    if (row < n_pad) {
      const double4 lv = *reinterpret_cast<const double4*>(Lb + row * ld + static_cast<int64_t>(k) * TILE + 4 * lane);
      double s = fma(lv.x, xv.x, fma(lv.y, xv.y, fma(lv.z, xv.z, lv.w * xv.w)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) rb[row] -= s;
    }
  }
}
// x (n_pad per batch entry) <- L^-1 r for every batch entry; r is used as scratch
void launch_trsv_l(const double* L, int64_t ld, int64_t l_bs, const double* Dinv, int64_t d_bs, int64_t n_pad,
                   double* r, int64_t r_bs, double* x, int64_t x_bs, int batch, cudaStream_t st, int k_begin) {
  const int nt = static_cast<int>(n_pad / TILE);
  for (int k = k_begin; k < nt; ++k) {
    const int64_t below = n_pad - static_cast<int64_t>(k + 1) * TILE;
    // batched sweeps have CTAs to spare: four row blocks per CTA quarter the re-reads of W
    const int groups = (batch >= 64 && below >= 512) ? 8 : ((batch >= 16 && below >= 256) ? 4 : 1);
    dim3 grid(static_cast<unsigned>(below > 0 ? (below + 64 * groups - 1) / (64 * groups) : 1), batch);
    launch_chain(trsv_l_step_kernel, grid, dim3(256), 0, st, g_pdl != 0, L, ld, l_bs, Dinv, d_bs, k, n_pad, r, r_bs, x, x_bs,
                 groups);
  }
}

// The same forward-substitution step for up to 16 right-hand sides that share L (rows of a thin appended block,
// grow.cu): the 64 x 128 slice of L is read ONCE per CTA and applied to all of them.
//   r, x: m rows with pitch `pitch` (right-hand side i at r + i * pitch); x is written, r is scratch.
constexpr int TRSV_MULTI_MAX = 16;
__global__ void __launch_bounds__(256) trsv_l_multi_step_kernel(const double* __restrict__ L, int64_t ld,
                                                                   const double* __restrict__ Dinv, int k, int64_t n_pad,
                                                                   double* __restrict__ r, double* __restrict__ x,
                                                                   int64_t pitch, int m, int groups) {
  __shared__ double rk[TRSV_MULTI_MAX][TILE];
  __shared__ __align__(32) double xk[TRSV_MULTI_MAX][TILE];
  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const int row = t & 127, h = t >> 7;
  pdl_trigger();
  pdl_wait();
  for (int e = t; e < m * TILE; e += 256) rk[e >> 7][e & 127] = r[(e >> 7) * pitch + k * TILE + (e & 127)];
  __syncthreads();
  {
    // x_k[i][row] = sum_c W[row][c] r_k[i][c]: thread (row, half of the columns); W is read once for all m
    const double* W = Dinv + static_cast<int64_t>(k) * TILE * TILE + row * TILE + 64 * h;
    double acc[TRSV_MULTI_MAX];
#pragma unroll
    for (int i = 0; i < TRSV_MULTI_MAX; ++i) acc[i] = 0.0;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 16) {
      double w[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) w[c] = W[c0 + c];
#pragma unroll
      for (int i = 0; i < TRSV_MULTI_MAX; ++i) {
        if (i < m) {
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            a0 = fma(w[c], rk[i][64 * h + c0 + c], a0);
            a1 = fma(w[c + 1], rk[i][64 * h + c0 + c + 1], a1);
          }
          acc[i] += a0 + a1;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < TRSV_MULTI_MAX; ++i)
      if (i < m && h == 0) xk[i][row] = acc[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TRSV_MULTI_MAX; ++i)
      if (i < m && h == 1) xk[i][row] += acc[i];
  }
  __syncthreads();
  if (blockIdx.x == 0)
    for (int e = t; e < m * TILE; e += 256) x[(e >> 7) * pitch + k * TILE + (e & 127)] = xk[e >> 7][e & 127];
  // rows below the tile: one warp per row, 4 consecutive doubles per lane; `groups` blocks of 64 rows per CTA
  for (int gi = 0; gi < groups; ++gi) {
    double4 lv[8];
    int64_t rows[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {                        // eight independent row loads in flight per warp
      rows[q] = static_cast<int64_t>(k + 1) * TILE + (static_cast<int64_t>(blockIdx.x) * groups + gi) * 64 + warp * 8 + q;
      lv[q] = rows[q] < n_pad ? *reinterpret_cast<const double4*>(L + rows[q] * ld + static_cast<int64_t>(k) * TILE + 4 * lane)
                              : make_double4(0.0, 0.0, 0.0, 0.0);
    }
    for (int i = 0; i < m; ++i) {
      const double4 xv = *reinterpret_cast<const double4*>(&xk[i][4 * lane]);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        double s = fma(lv[q].x, xv.x, fma(lv[q].y, xv.y, fma(lv[q].z, xv.z, lv[q].w * xv.w)));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0 && rows[q] < n_pad) r[i * pitch + rows[q]] -= s;
      }
    }
  }
}
void launch_trsv_l_multi(const double* L, int64_t ld, const double* Dinv, int64_t n_pad, double* r, double* x,
                         int64_t pitch, int m, cudaStream_t st) {
  GPB_REQUIRE(m >= 1 && m <= TRSV_MULTI_MAX, "trsv_l_multi: between 1 and 16 right-hand sides");
  const int nt = static_cast<int>(n_pad / TILE);
  for (int k = 0; k < nt; ++k) {
    const int64_t below = n_pad - static_cast<int64_t>(k + 1) * TILE;
    const int groups = 1;                                // (more rows per CTA measured no faster)
    dim3 grid(static_cast<unsigned>(below > 0 ? (below + 64 * groups - 1) / (64 * groups) : 1));
    launch_chain(trsv_l_multi_step_kernel, grid, dim3(256), 0, st, g_pdl != 0, L, ld, Dinv, k, n_pad, r, x, pitch, m, groups);
  }
}

__global__ void fill_kernel(double* p, int64_t n, double v) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}
void launch_fill(double* p, int64_t n, double v, cudaStream_t st) {
  if (n <= 0) return;
  const int blocks = static_cast<int>((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  fill_kernel<<<blocks, 256, 0, st>>>(p, n, v);
  GPB_CUDA(cudaGetLastError());
}

__global__ void copy_sub_mean_kernel(double* dst, const double* src, int64_t n, double mean) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = src[i] - mean;
}
void launch_copy_sub_mean(double* dst, const double* src, int64_t n, double mean, cudaStream_t st) {
  if (n <= 0) return;
  const int blocks = static_cast<int>((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  copy_sub_mean_kernel<<<blocks, 256, 0, st>>>(dst, src, n, mean);
  GPB_CUDA(cudaGetLastError());
}

}  // namespace gpb
