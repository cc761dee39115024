// laplace.cu - Laplace / Newton mode finding on the device.
//
//  * preference GP (PreferenceGaussianProcess.calc_laplace, GPpref.py:112-157, PrefProbit
//    GPpref.py:46-94) including the reference's quirks (see gpb200.h);
//  * binary classification (GPc.py intent; Rasmussen & Williams Alg. 3.1 / 3.2).
//
// Per Newton iteration the reference inverts a dense n x n matrix on the host
// (np.linalg.inv, GPpref.py:143) and builds W with a Python loop over the pairs
// (GPpref.py:82-87).  Here one iteration is: a fused per-pair kernel (z, Phi, N, gradient and
// Hessian weights), a per-item kernel that assembles G = K^-1 + W and the right-hand side from a
// CSR view of the comparison graph (deterministic, same per-cell summation order as the
// reference's loop), the blocked DMMA Cholesky of G with the right-hand side riding along as an
// appended row, a blocked backward substitution, and one reduction kernel for the objective and
// max|f_new - f|.  The host only reads back 16 bytes per iteration to decide on convergence.
#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/gpb200.h"
#include <chrono>
#include <cstdlib>

#include "gpb_context.cuh"

using namespace gpb;

namespace {

constexpr double LOG_2PI = 1.8378770664093453;
constexpr double INV_SQRT_2PI = 0.3989422804014327;
constexpr double SQRT_2_OVER_PI = 0.7978845608028654;
constexpr double INV_SQRT_2 = 0.7071067811865476;

// ---------------------------------------------------------------------------------------
// generic helpers
// ---------------------------------------------------------------------------------------
template <int NT>
__device__ __forceinline__ double cta_sum(double v, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
#pragma unroll
  for (int w = NT / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}
template <int NT>
__device__ __forceinline__ double cta_max(double v, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
#pragma unroll
  for (int w = NT / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + w]);
    __syncthreads();
  }
  const double r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(512) sum_log_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
  __shared__ double sh[512];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 512) s += log(v[i]);
  const double r = cta_sum<512>(s, sh);
  if (threadIdx.x == 0) out[0] = r;
}

// full symmetric copy of a matrix whose lower 64-tiles (and full diagonal tiles) are valid
__global__ void __launch_bounds__(256) mirror_lower_kernel(const double* __restrict__ src, int64_t ld_s,
                                                           double* __restrict__ dst, int64_t ld_d, int64_t nt64) {
  __shared__ double tt[64][65];
  const int64_t ti = blockIdx.x / nt64, tj = blockIdx.x % nt64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (ti >= tj) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int64_t r = ti * 64 + ty + 16 * a, cc = tj * 64 + tx + 16 * c;
        dst[r * ld_d + cc] = src[r * ld_s + cc];
      }
  } else {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tt[ty + 16 * a][tx + 16 * c] = src[(tj * 64 + ty + 16 * a) * ld_s + ti * 64 + tx + 16 * c];
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c)
        dst[(ti * 64 + ty + 16 * a) * ld_d + tj * 64 + tx + 16 * c] = tt[tx + 16 * c][ty + 16 * a];
  }
}

// lower triangle (row by row) of dst <- rs[r] * src * rs[c] + (r == c ? dadd : 0); rs == nullptr: plain copy
__global__ void __launch_bounds__(256) scale_copy_lower_kernel(const double* __restrict__ src, int64_t ld_s,
                                                               double* __restrict__ dst, int64_t ld_d,
                                                               const double* __restrict__ rs, double dadd) {
  const int64_t r = blockIdx.x;
  const double sr = rs ? rs[r] : 1.0;
  for (int64_t c = threadIdx.x; c <= r; c += 256) {
    double v = src[r * ld_s + c];
    if (rs) v = sr * v * rs[c];
    if (c == r) v += dadd;
    dst[r * ld_d + c] = v;
  }
}

// One step of the blocked backward substitution  L^T x = r  with the inverted diagonal tiles:
//   x_k = W_k^T r_k ;  r_c -= sum_{rows of tile k} L[row][c] x_k[row]   for every column c left of tile k.
__global__ void __launch_bounds__(256) trsv_lt_step_kernel(const double* __restrict__ L, int64_t ld,
                                                           const double* __restrict__ Dinv, int k,
                                                           double* __restrict__ r, double* __restrict__ x) {
  // written for memory-level parallelism: every load below is independent of the running sums
  __shared__ double rk[TILE];
  __shared__ double xk[TILE];
  __shared__ double part[256];
  const int t = threadIdx.x;
  pdl_trigger();
  pdl_wait();                                    // r comes from the previous step
  if (t < TILE) rk[t] = r[k * TILE + t];
  __syncthreads();
  {
    // x_k = W_k^T r_k on all 256 threads: column c, row half h (rows above the diagonal of W hold zeros)
    const int c = t & 127, h = t >> 7;
    const double* W = Dinv + static_cast<int64_t>(k) * TILE * TILE + c;
    double acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = 0.0;
#pragma unroll
    for (int i = 0; i < 64; i += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int row = 64 * h + i + u;
        acc[u] = fma(W[row * TILE], rk[row], acc[u]);
      }
    }
    part[t] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  }
  __syncthreads();
  if (t < TILE) {
    const double s = part[t] + part[t + 128];
    xk[t] = s;
    if (blockIdx.x == 0) x[k * TILE + t] = s;
  }
  __syncthreads();
  // r_c -= L[tile k rows][c] . x_k for 64 columns per CTA, the 128 rows cut into 4 quarters
  const int cl = t & 63, q = t >> 6;
  const int64_t c = static_cast<int64_t>(blockIdx.x) * 64 + cl;
  const bool on = c < static_cast<int64_t>(k) * TILE;
  double s = 0.0;
  if (on) {
    const double* Lk = L + (static_cast<int64_t>(k) * TILE + 32 * q) * ld + c;
    double a[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] = fma(Lk[(i + u) * ld], xk[32 * q + i + u], a[u]);
    }
    s = (a[0] + a[1]) + (a[2] + a[3]);
  }
  part[t] = s;
  __syncthreads();
  if (q == 0 && on) r[c] -= (part[cl] + part[cl + 64]) + (part[cl + 128] + part[cl + 192]);
}

void trsv_lt(gpb_handle* h, const FactorMat& m, double* r, double* x) {
  const int nt = static_cast<int>(m.n_pad / TILE);
  for (int k = nt - 1; k >= 0; --k) {
    const int blocks = k == 0 ? 1 : 2 * k;
    launch_chain(trsv_lt_step_kernel, dim3(blocks), dim3(256), 0, h->s0, g_pdl != 0, m.A, m.ld, m.Dinv, k, r, x);
    ++h->launches;
  }
}

// FactorMat over h->A with the (2 np + 128)-row layout shared with grad.cu
FactorMat laplace_mat(gpb_handle* h, int64_t rows_total) {
  const int64_t np = h->n_pad;
  const int64_t rows_alloc = 3 * np + TILE;     // factor | appended tile row | scratch S1 (np rows) | scratch S2 (np rows)
  FactorMat m;
  m.ld = np; m.n_pad = np; m.rows_total = rows_total; m.batch = 1;
  m.batch_stride = rows_alloc * np;
  h->A.ensure(static_cast<size_t>(m.batch_stride) * 8);
  m.A = h->A.as<double>();
  m.dinv_bs = np * TILE;
  h->Dinv.ensure(static_cast<size_t>(m.dinv_bs) * 8);
  m.Dinv = h->Dinv.as<double>();
  m.diag_bs = np;
  h->diag.ensure(static_cast<size_t>(np) * 8);
  m.diag = h->diag.as<double>();
  h->info.ensure(64);
  m.info = h->info.as<int>();
  finalize_factor_mat(m);
  make_tile_maps(&m.mapA, m.A, np, rows_alloc, 1, np, m.batch_stride);
  return m;
}

// K (+ jitter on the diagonal) from the training inputs: lower tiles into dst (mode 1) or full (mode 0)
void build_kernel_matrix(gpb_handle* h, const double* ell_dev, const double* hyp2_dev, double* dst, int64_t ld, int mode) {
  const int64_t np = h->n_pad;
  h->XsT.ensure(static_cast<size_t>(h->d) * np * 8);
  h->sq.ensure(static_cast<size_t>(np) * 8);
  launch_se_prep(h->X.as<double>(), h->n, h->d, ell_dev, h->XsT.as<double>(), np, h->sq.as<double>(), 1, 0, 0, 0, h->s0);
  SeArgs a{};
  a.rT = a.cT = h->XsT.as<double>(); a.r_ld = a.c_ld = np;
  a.r_sq = a.c_sq = h->sq.as<double>();
  a.n_rows_valid = a.n_cols_valid = h->n;
  a.d = h->d; a.out = dst; a.ld = ld; a.rows_pad = a.cols_pad = np;
  a.hyp_dev = hyp2_dev; a.mode = mode; a.clip = 1;        // GPy RBF semantics (GPpref.py:122)
  launch_se_build(a, 1, h->s0);
  h->launches += 2;
}

int read_info(gpb_handle* h, const FactorMat& m) {
  double* host = h->pinned(64);
  GPB_CUDA(cudaMemcpyAsync(host, m.info, 4, cudaMemcpyDeviceToHost, h->s0));
  GPB_CUDA(cudaStreamSynchronize(h->s0));
  return *reinterpret_cast<int*>(host);
}

// upload [ell..., variance, jitter]; returns device pointers (ell, hyp2)
void upload_kernel_params(gpb_handle* h, const double* khyp, double jitter, const double** ell, const double** hyp2) {
  const int d = h->d;
  double* host = h->pinned((d + 2) * 8);
  for (int k = 0; k < d; ++k) host[k] = khyp[k];
  host[d] = khyp[d];
  host[d + 1] = jitter;
  h->params.ensure((d + 2) * 8);
  GPB_CUDA(cudaMemcpyAsync(h->params.p, host, (d + 2) * 8, cudaMemcpyHostToDevice, h->s0));
  GPB_CUDA(cudaStreamSynchronize(h->s0));
  *ell = h->params.as<double>();
  *hyp2 = h->params.as<double>() + d;
}

// ---------------------------------------------------------------------------------------
// preference likelihood (PrefProbit, GPpref.py:46-94)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double pref_hazard(double z) {
  // N(z)/Phi(z) as the reference forms it (GPpref.py:71-72,76); where Phi underflows (z < -37,
  // the reference produces 0/0 = nan) the erfcx form keeps the iteration finite.
  const double phi = normcdf(z);
  if (phi > 1e-300) return exp(-0.5 * z * z) * INV_SQRT_2PI / phi;
  return SQRT_2_OVER_PI / erfcx(-z * INV_SQRT_2);
}

// per pair: d_k = y isqrt2sig N/Phi (GPpref.py:76), w_k = i2var (z N/Phi + (N/Phi)^2) = -inner (GPpref.py:80)
__global__ void pref_pair_kernel(const int64_t* __restrict__ uvi, const double* __restrict__ y, int64_t P,
                                 const double* __restrict__ f, double isqrt2sig, double i2var,
                                 double* __restrict__ dk, double* __restrict__ wk) {
  const int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (k >= P) return;
  const int64_t u = uvi[2 * k], v = uvi[2 * k + 1];
  const double z = y[k] * (isqrt2sig * (f[v] - f[u]));       // GPpref.py:56-58
  const double r = pref_hazard(z);
  dk[k] = y[k] * isqrt2sig * r;
  wk[k] = i2var * (z * r + r * r);
}

struct PrefGraphDev {
  const int64_t* row_ptr;      // n + 1
  const int32_t* other;        // per row: the other item of each incident pair, sorted by (other, pair)
  const int32_t* pair_by_other;
  const int32_t* is_u;         // 1: the row item is u (column 0) of that pair
  const int32_t* pair_sorted;  // per row: incident pairs in ascending pair order (diagonal cell)
  const int64_t* ku;           // last pair with u == i, or -1 (GPpref.py:77 last write wins)
  const int64_t* kv;           // last pair with v == i, or -1 (GPpref.py:78)
};

// per item i: gradient g_i, row i of G = K^-1 + W (lower part), b_i = (W f)_i + g_i.
// Every cell of W is accumulated in pair order like the reference's loop (GPpref.py:82-87) and only
// then added to K^-1 (GPpref.py:142).
__global__ void pref_row_kernel(int64_t n, const PrefGraphDev gr, const double* __restrict__ dk,
                                const double* __restrict__ wk, const double* __restrict__ f, int grad_mode,
                                double* __restrict__ G, int64_t ld, double* __restrict__ b,
                                double* __restrict__ g_out, double* __restrict__ Wdense) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  double g = 0.0;
  if (grad_mode == 0) {
    if (gr.ku[i] >= 0) g = 0.0 - dk[gr.ku[i]];     // column-0 pass: g[u] = 0 + (-d)
    if (gr.kv[i] >= 0) g = g + dk[gr.kv[i]];       // column-1 pass adds to whatever pass 0 left
  }
  const int64_t e0 = gr.row_ptr[i], e1 = gr.row_ptr[i + 1];
  // One thread walks the ~2P/n incident pairs of its item.  Done entry by entry this is a chain of four dependent
  // global loads per entry (index, pair, value, cell: 3 us each, 49 us per Newton iteration at C4 size on 32 SMs);
  // the loads of eight entries are issued together instead, and only the sums stay sequential (same order).
  constexpr int CH = 8;
  double wf = 0.0;
  int32_t cur_j = -1;
  double wij = 0.0, f_cur = 0.0, g_cur = 0.0;
  auto close_group = [&]() {
    if (cur_j < 0) return;
    wf = fma(wij, f_cur, wf);
    if (G && cur_j < i) G[i * ld + cur_j] = g_cur + wij;
    if (Wdense) Wdense[i * n + cur_j] = wij;
  };
  for (int64_t base = e0; base < e1; base += CH) {
    int32_t oj[CH], pk[CH];
    double wv[CH], fv[CH], gv[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const bool on = base + q < e1;
      oj[q] = on ? gr.other[base + q] : -1;
      pk[q] = on ? gr.pair_by_other[base + q] : 0;
    }
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const bool on = oj[q] >= 0;
      wv[q] = on ? wk[pk[q]] : 0.0;
      fv[q] = on ? f[oj[q]] : 0.0;
      gv[q] = (on && G && oj[q] < i) ? G[i * ld + oj[q]] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      if (oj[q] >= 0) {
        if (oj[q] != cur_j) {
          close_group();
          cur_j = oj[q]; wij = 0.0; f_cur = fv[q]; g_cur = gv[q];
        }
        wij -= wv[q];                                // W[xi, yi] -= -ddpy_df (GPpref.py:86-87)
      }
    }
  }
  close_group();
  double wii = 0.0;
  for (int64_t base = e0; base < e1; base += CH) {
    double wv[CH];
#pragma unroll
    for (int q = 0; q < CH; ++q) wv[q] = base + q < e1 ? wk[gr.pair_sorted[base + q]] : 0.0;
#pragma unroll
    for (int q = 0; q < CH; ++q)
      if (base + q < e1) wii += wv[q];               // W[xi, xi] -= ddpy_df (GPpref.py:84-85)
  }
  if (grad_mode == 1) {
    for (int64_t q = e0; q < e1; ++q) g += gr.is_u[q] ? -dk[gr.pair_by_other[q]] : dk[gr.pair_by_other[q]];
  }
  wf = fma(wii, f[i], wf);
  if (G) G[i * ld + i] += wii;
  if (Wdense) Wdense[i * n + i] = wii;
  if (b) b[i] = wf + g;
  if (g_out) g_out[i] = g;
}

// ---------------------------------------------------------------------------------------
// The Newton loops run on the device (GPpref.py:140-155 tests convergence on the host and prints once per
// iteration).  One iteration is captured into the body of a CUDA-graph WHILE node; the last kernel of the
// iteration (the finish kernel, thread 0) appends (f_error, objective) to a device trace, counts the iteration,
// looks at the factorisation's info word and decides whether the loop goes on: cudaGraphSetConditional.  The
// host launches the graph once and reads counter, status and trace once when it is done - nothing between
// iterations crosses PCIe and the stream never drains.
// ---------------------------------------------------------------------------------------
struct LoopCtl {
  int it;          // iterations completed
  int status;      // 0 running / converged, 2: the factorisation inside an iteration failed (pivot holds info)
  int pivot;
  int max_iter;
  double delta_f;
};
__device__ __forceinline__ void loop_step(LoopCtl* ctl, double* trace, const int* info, cudaGraphConditionalHandle hc,
                                          double f_error, double objective) {
  const int it = ctl->it;
  trace[2 * it] = f_error;
  trace[2 * it + 1] = objective;
  ctl->it = it + 1;
  const int piv = *info;
  if (piv != 0) { ctl->status = 2; ctl->pivot = piv; }
  // GPpref.py:140: while f_error > delta_f (a NaN error ends the loop, as it does on the host)
  const bool go = piv == 0 && (f_error > ctl->delta_f) && (it + 1 < ctl->max_iter);
  if (hc) cudaGraphSetConditional(hc, go ? 1u : 0u);
  else ctl->status = (ctl->status == 2) ? 2 : (go ? 0 : 1);      // diagnostic host-driven mode (GPB_HOST_LOOP): 1 = stop
}

// lml = sum log Phi(z(f_new)) - 0.5 f_new' iK f_new - 0.5 logdetK - n/2 log(2 pi)   (GPpref.py:90-94)
// f_error = max |f_new - f| (GPpref.py:151-152); then f <- f_new (GPpref.py:155)
// Two stages: the P log-Phi terms are ~100 FP64 operations each - on ONE SM (round 1) that was 26 us of its FP64
// pipe per Newton iteration; now PREF_PARTS CTAs reduce their slices into partials, and the one-CTA finish kernel
// adds the partials in a fixed order (deterministic) and takes the loop decision.
constexpr int PREF_PARTS = 64;
__global__ void __launch_bounds__(256) pref_terms_kernel(const int64_t* __restrict__ uvi, const double* __restrict__ y,
                                                         int64_t P, int64_t n, double isqrt2sig,
                                                         const double* __restrict__ f_new, const double* __restrict__ f,
                                                         const double* __restrict__ t, double* __restrict__ part) {
  __shared__ double sh[256];
  const int64_t stride = static_cast<int64_t>(gridDim.x) * 256;
  double s = 0.0;
  for (int64_t k = blockIdx.x * 256 + threadIdx.x; k < P; k += stride) {
    const double z = y[k] * (isqrt2sig * (f_new[uvi[2 * k + 1]] - f_new[uvi[2 * k]]));
    s += log(normcdf(z));
  }
  const double slog = cta_sum<256>(s, sh);
  double q = 0.0, mx = 0.0;
  for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += stride) {
    q = fma(f_new[i], t[i], q);
    mx = fmax(mx, fabs(f_new[i] - f[i]));
  }
  const double qs = cta_sum<256>(q, sh);
  const double m = cta_max<256>(mx, sh);
  if (threadIdx.x == 0) {
    part[blockIdx.x] = slog;
    part[PREF_PARTS + blockIdx.x] = qs;
    part[2 * PREF_PARTS + blockIdx.x] = m;
  }
}
__global__ void __launch_bounds__(256) pref_finish_kernel(int64_t n, const double* __restrict__ f_new, double* __restrict__ f,
                                                          const double* __restrict__ part, const double* __restrict__ logdet,
                                                          double* __restrict__ out2, LoopCtl* ctl = nullptr,
                                                          double* trace = nullptr, const int* info = nullptr,
                                                          cudaGraphConditionalHandle hc = 0) {
  for (int64_t i = threadIdx.x; i < n; i += 256) f[i] = f_new[i];
  if (threadIdx.x == 0) {
    double slog = 0.0, qs = 0.0, m = 0.0;
    for (int b = 0; b < PREF_PARTS; ++b) {
      slog += part[b];
      qs += part[PREF_PARTS + b];
      m = fmax(m, part[2 * PREF_PARTS + b]);
    }
    const double lml = slog - 0.5 * qs - 0.5 * logdet[0] - 0.5 * static_cast<double>(n) * LOG_2PI;
    out2[0] = m;
    out2[1] = lml;
    if (ctl) loop_step(ctl, trace, info, hc, m, lml);
  }
}

struct PrefGraphHost {
  std::vector<int64_t> row_ptr, ku, kv;
  std::vector<int32_t> other, pair_by_other, is_u, pair_sorted;
};

PrefGraphHost build_pref_graph(const int64_t* uvi, int64_t P, int64_t n) {
  PrefGraphHost g;
  g.row_ptr.assign(n + 1, 0);
  g.ku.assign(n, -1);
  g.kv.assign(n, -1);
  for (int64_t k = 0; k < P; ++k) {
    const int64_t u = uvi[2 * k], v = uvi[2 * k + 1];
    if (u < 0 || u >= n || v < 0 || v >= n) throw Error{"uvi index out of range"};
    g.ku[u] = k;                       // later pairs overwrite earlier ones
    g.kv[v] = k;
    if (u != v) { ++g.row_ptr[u + 1]; ++g.row_ptr[v + 1]; }     // u == v contributes exactly zero to W
  }
  for (int64_t i = 0; i < n; ++i) g.row_ptr[i + 1] += g.row_ptr[i];
  const int64_t nnz = g.row_ptr[n];
  g.other.resize(nnz); g.pair_by_other.resize(nnz); g.is_u.resize(nnz); g.pair_sorted.resize(nnz);
  std::vector<int64_t> fill(g.row_ptr.begin(), g.row_ptr.end() - 1);
  struct Ent { int32_t other, pair, is_u; };
  std::vector<Ent> ent(nnz);
  for (int64_t k = 0; k < P; ++k) {
    const int64_t u = uvi[2 * k], v = uvi[2 * k + 1];
    if (u == v) continue;
    ent[fill[u]++] = Ent{static_cast<int32_t>(v), static_cast<int32_t>(k), 1};
    ent[fill[v]++] = Ent{static_cast<int32_t>(u), static_cast<int32_t>(k), 0};
  }
  for (int64_t i = 0; i < n; ++i) {
    const int64_t e0 = g.row_ptr[i], e1 = g.row_ptr[i + 1];
    for (int64_t e = e0; e < e1; ++e) g.pair_sorted[e] = ent[e].pair;      // filled in ascending pair order
    std::stable_sort(ent.begin() + e0, ent.begin() + e1, [](const Ent& a, const Ent& b) { return a.other < b.other; });
    for (int64_t e = e0; e < e1; ++e) {
      g.other[e] = ent[e].other;
      g.pair_by_other[e] = ent[e].pair;
      g.is_u[e] = ent[e].is_u;
    }
  }
  return g;
}

// device copy of the graph + pairs, carved out of one buffer
struct PrefDev {
  PrefGraphDev gr;
  const int64_t* uvi;
  const double* y;
  double *dk, *wk;
};

PrefDev upload_pref(gpb_handle* h, const PrefGraphHost& g, const int64_t* uvi, const double* y, int64_t P, int64_t n) {
  const int64_t nnz = g.row_ptr[n];
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
  const size_t o_rp = take((n + 1) * 8), o_ku = take(n * 8), o_kv = take(n * 8), o_uvi = take(P * 16),
               o_y = take(P * 8), o_dk = take(P * 8), o_wk = take(P * 8), o_ot = take(nnz * 4 + 4),
               o_pb = take(nnz * 4 + 4), o_iu = take(nnz * 4 + 4), o_ps = take(nnz * 4 + 4);
  h->aux1.ensure(off);
  char* base = h->aux1.as<char>();
  auto up = [&](size_t o, const void* src, size_t bytes) {
    if (bytes) GPB_CUDA(cudaMemcpyAsync(base + o, src, bytes, cudaMemcpyHostToDevice, h->s0));
  };
  up(o_rp, g.row_ptr.data(), (n + 1) * 8); up(o_ku, g.ku.data(), n * 8); up(o_kv, g.kv.data(), n * 8);
  up(o_uvi, uvi, P * 16); up(o_y, y, P * 8);
  up(o_ot, g.other.data(), nnz * 4); up(o_pb, g.pair_by_other.data(), nnz * 4);
  up(o_iu, g.is_u.data(), nnz * 4); up(o_ps, g.pair_sorted.data(), nnz * 4);
  GPB_CUDA(cudaStreamSynchronize(h->s0));
  PrefDev d;
  d.gr.row_ptr = reinterpret_cast<const int64_t*>(base + o_rp);
  d.gr.ku = reinterpret_cast<const int64_t*>(base + o_ku);
  d.gr.kv = reinterpret_cast<const int64_t*>(base + o_kv);
  d.gr.other = reinterpret_cast<const int32_t*>(base + o_ot);
  d.gr.pair_by_other = reinterpret_cast<const int32_t*>(base + o_pb);
  d.gr.is_u = reinterpret_cast<const int32_t*>(base + o_iu);
  d.gr.pair_sorted = reinterpret_cast<const int32_t*>(base + o_ps);
  d.uvi = reinterpret_cast<const int64_t*>(base + o_uvi);
  d.y = reinterpret_cast<const double*>(base + o_y);
  d.dk = reinterpret_cast<double*>(base + o_dk);
  d.wk = reinterpret_cast<double*>(base + o_wk);
  return d;
}

// ---------------------------------------------------------------------------------------
// classification likelihood (GPc.py:4-21; R&W eq. 3.15 / 3.16)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void gpc_terms(int link, double y, double f, double& lp, double& g, double& W) {
  if (link == 0) {                                  // probit: log Phi(y f)  (GPc.py:5-6)
    const double x = y * f;
    const double ex = erfcx(-x * INV_SQRT_2);       // Phi(x) = 0.5 erfcx(-x/sqrt2) exp(-x^2/2)
    const double r = SQRT_2_OVER_PI / ex;           // N(f)/Phi(yf), finite for every x
    lp = (x < 0.0) ? log(0.5 * ex) - 0.5 * x * x : log(normcdf(x));
    g = y * r;
    W = r * r + x * r;
  } else {                                          // logistic (GPc.py:13-14)
    const double x = y * f;
    lp = (x > 0.0) ? -log1p(exp(-x)) : x - log1p(exp(x));
    const double pi = 1.0 / (1.0 + exp(-f));
    g = 0.5 * (y + 1.0) - pi;
    W = pi * (1.0 - pi);
  }
}

// b = W f + g ; sW = sqrt(W) ; pads get W = 0
__global__ void gpc_terms_kernel(int link, const double* __restrict__ y, const double* __restrict__ f, int64_t n,
                                 int64_t n_pad, double* __restrict__ sW, double* __restrict__ b, double* __restrict__ g_out) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_pad) return;
  if (i >= n) { sW[i] = 0.0; b[i] = 0.0; if (g_out) g_out[i] = 0.0; return; }
  double lp, g, W;
  gpc_terms(link, y[i], f[i], lp, g, W);
  sW[i] = sqrt(W);
  b[i] = fma(W, f[i], g);
  if (g_out) g_out[i] = g;
}
__global__ void vec_mul_kernel(double* __restrict__ dst, const double* __restrict__ a, const double* __restrict__ b, int64_t n) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = a[i] * b[i];
}
// a = b - sW * t
__global__ void gpc_a_kernel(double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ sW,
                             const double* __restrict__ t, int64_t n) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) a[i] = b[i] - sW[i] * t[i];
}
// out2 = (max|f_new - f|, -0.5 a'f_new + sum log p(y|f_new)); f <- f_new
__global__ void __launch_bounds__(1024) gpc_finish_kernel(int link, const double* __restrict__ y, int64_t n,
                                                          const double* __restrict__ a, const double* __restrict__ f_new,
                                                          double* __restrict__ f, double* __restrict__ out2, LoopCtl* ctl,
                                                          double* trace, const int* info, cudaGraphConditionalHandle hc) {
  __shared__ double sh[1024];
  double q = 0.0, mx = 0.0, sl = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    double lp, g, W;
    gpc_terms(link, y[i], f_new[i], lp, g, W);
    sl += lp;
    q = fma(a[i], f_new[i], q);
    mx = fmax(mx, fabs(f_new[i] - f[i]));
    f[i] = f_new[i];
  }
  const double qs = cta_sum<1024>(q, sh);
  const double ls = cta_sum<1024>(sl, sh);
  const double m = cta_max<1024>(mx, sh);
  if (threadIdx.x == 0) {
    out2[0] = m;
    out2[1] = -0.5 * qs + ls;
    if (ctl) loop_step(ctl, trace, info, hc, m, -0.5 * qs + ls);
  }
}
// rows[i][j] *= s[j]
__global__ void scale_cols_kernel(double* __restrict__ rows, int64_t ld, int64_t ncols, const double* __restrict__ s) {
  double* r = rows + blockIdx.x * ld;
  for (int64_t j = threadIdx.x; j < ncols; j += blockDim.x) r[j] *= s[j];
}
// var_i = sf2 - |row_i|^2 ; prob_i from (mu_i, var_i)
__global__ void __launch_bounds__(256) gpc_predict_finish_kernel(const double* __restrict__ rows, int64_t ld, int64_t ncols,
                                                                 double sf2, int link, const double* __restrict__ mu,
                                                                 double* __restrict__ var, double* __restrict__ prob) {
  __shared__ double sh[256];
  const double* r = rows + blockIdx.x * ld;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < ncols; j += 256) s = fma(r[j], r[j], s);
  const double ss = cta_sum<256>(s, sh);
  if (threadIdx.x == 0) {
    const double v = sf2 - ss;
    var[blockIdx.x] = v;
    const double m = mu[blockIdx.x];
    prob[blockIdx.x] = (link == 0) ? normcdf(m / sqrt(1.0 + v))                                   // R&W eq. 3.82
                                   : 1.0 / (1.0 + exp(-m / sqrt(1.0 + 0.39269908169872414 * v)));   // MacKay, pi/8
  }
}

// ---- state kept by gpb_pref_laplace for gpb_pref_evidence / gpb_pref_predict ------------------------------
struct PrefState {
  int64_t n = 0, P = 0;
  double sigma = 1.0, eps = 0.0, lml_ref = 0.0, half_logdet_k = 0.0;
  int grad_mode = 0;
  bool factored = false;          // chol(K^-1 + W(f_hat)) is in the work space
  double sum_log_g = 0.0;         // sum log diag of that factor
  std::vector<double> khyp;
  PrefDev pd;
};
void pref_state_free(void* p) { delete static_cast<PrefState*>(p); }

// rows <- rows_b - rows (difference of two covariance row blocks); one CTA per row
__global__ void rows_sub_kernel(double* __restrict__ rows, const double* __restrict__ rows_b, int64_t ld, int64_t ncols) {
  double* r = rows + blockIdx.x * ld;
  const double* q = rows_b + blockIdx.x * ld;
  for (int64_t j = threadIdx.x; j < ncols; j += blockDim.x) r[j] = q[j] - r[j];
}
// k** of the quantity predicted: sf2 for one item, k(a,a) + k(b,b) - 2 k(a,b) for a difference (scaled points, pitch ld_t)
__global__ void pref_kss_kernel(const double* __restrict__ zaT, const double* __restrict__ zbT, int64_t ld_t, int d, int64_t m,
                                double sf2, double* __restrict__ kss) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= m) return;
  if (!zbT) { kss[i] = sf2; return; }
  double r2 = 0.0;
  for (int k = 0; k < d; ++k) {
    const double t = zaT[k * ld_t + i] - zbT[k * ld_t + i];
    r2 = fma(t, t, r2);
  }
  kss[i] = 2.0 * sf2 - 2.0 * sf2 * exp(-0.5 * r2);
}
// q1_i = R_i . T_i (two row blocks with the same pitch)
__global__ void __launch_bounds__(256) rows_dot2_kernel(const double* __restrict__ R, const double* __restrict__ T, int64_t ld,
                                                        int64_t ncols, double* __restrict__ out) {
  __shared__ double sh[256];
  const double* r = R + blockIdx.x * ld;
  const double* t = T + blockIdx.x * ld;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < ncols; j += 256) s = fma(r[j], t[j], s);
  const double ss = cta_sum<256>(s, sh);
  if (threadIdx.x == 0) out[blockIdx.x] = ss;
}
// var_i = kss_i - q1_i + |V_i|^2 ; prob_i = Phi(mean_i / sqrt(2 sigma^2 + var_i)) for differences
__global__ void __launch_bounds__(256) pref_predict_finish_kernel(const double* __restrict__ V, int64_t ld, int64_t ncols,
                                                                  const double* __restrict__ kss, const double* __restrict__ q1,
                                                                  const double* __restrict__ mean, double two_sigma2,
                                                                  double* __restrict__ var, double* __restrict__ prob) {
  __shared__ double sh[256];
  const double* r = V + blockIdx.x * ld;
  double s = 0.0;
  for (int64_t j = threadIdx.x; j < ncols; j += 256) s = fma(r[j], r[j], s);
  const double ss = cta_sum<256>(s, sh);
  if (threadIdx.x == 0) {
    const double v = kss[blockIdx.x] - q1[blockIdx.x] + ss;
    var[blockIdx.x] = v;
    if (prob) prob[blockIdx.x] = normcdf(mean[blockIdx.x] / sqrt(two_sigma2 + fmax(v, 0.0)));
  }
}

// chol(K^-1 + W(f_hat)) into the work space (the loop leaves the factor of the PREVIOUS iterate there)
void pref_factor_at_mode(gpb_handle* h, PrefState* ps) {
  if (ps->factored) return;
  const int64_t n = ps->n, np = h->n_pad;
  FactorMat g = laplace_mat(h, np);
  const double* iK = h->aux2.as<double>();
  double* f = h->aux0.as<double>();
  const double isq = 1.0 / (ps->sigma * std::sqrt(2.0)), i2v = isq * isq;
  pref_pair_kernel<<<static_cast<unsigned>((ps->P + 255) / 256), 256, 0, h->s0>>>(ps->pd.uvi, ps->pd.y, ps->P, f, isq, i2v,
                                                                               ps->pd.dk, ps->pd.wk);
  scale_copy_lower_kernel<<<static_cast<unsigned>(np), 256, 0, h->s0>>>(iK, np, g.A, g.ld, nullptr, 0.0);
  pref_row_kernel<<<static_cast<unsigned>((n + 63) / 64), 64, 0, h->s0>>>(n, ps->pd.gr, ps->pd.dk, ps->pd.wk, f, 1, g.A,
                                                                            g.ld, nullptr, nullptr, nullptr);
  GPB_CUDA(cudaGetLastError());
  GPB_CUDA(cudaMemsetAsync(g.info, 0, 4, h->s0));
  chol_sweep(h, g, true);
  double* sc = h->scal.as<double>();
  sum_log_kernel<<<1, 512, 0, h->s0>>>(g.diag, np, sc + 8);
  GPB_CUDA(cudaGetLastError());
  h->launches += 4;
  double* host = h->pinned(64);
  GPB_CUDA(cudaMemcpyAsync(host, sc + 8, 8, cudaMemcpyDeviceToHost, h->s0));
  GPB_CUDA(cudaMemcpyAsync(host + 1, g.info, 4, cudaMemcpyDeviceToHost, h->s0));
  GPB_CUDA(cudaStreamSynchronize(h->s0));
  if (*reinterpret_cast<int*>(host + 1)) throw Error{"K^-1 + W is not positive definite at the mode"};
  ps->sum_log_g = host[0];
  ps->factored = true;
}

PrefState* pref_state_checked(gpb_handle* h) {
  PrefState* ps = static_cast<PrefState*>(h->pref_state);
  // the entry macro has already counted this call: the state is current iff nothing else ran in between
  GPB_REQUIRE(ps != nullptr && ps->n == h->n && h->state_epoch + 1 == h->ws_epoch,
              "no preference Laplace state on this handle: call gpb_pref_laplace first (any other call that uses "
              "the work space invalidates it)");
  return ps;
}

// Runs `iteration(handle)` - which enqueues ONE Newton iteration on h->s0 (side streams joined by events, as
// chol_sweep does) and ends with a finish kernel calling loop_step - as the body of a WHILE node until the device
// says stop.  Returns the number of kernel launches of one iteration.  ctl / trace: device; result read by the caller.
struct LoopResult { int it, status, pivot; int64_t launches_per_iteration; };
template <class F>
LoopResult run_device_loop_once(gpb_handle* h, LoopCtl* ctl_dev, double* trace_dev, double* trace_host, double delta_f,
                                int max_iter, F& iteration) {
  LoopCtl init{0, 0, 0, max_iter, delta_f};
  LoopCtl* hc_host = reinterpret_cast<LoopCtl*>(h->pinned(sizeof(LoopCtl) + 64));
  *hc_host = init;
  GPB_CUDA(cudaMemcpyAsync(ctl_dev, hc_host, sizeof(LoopCtl), cudaMemcpyHostToDevice, h->s0));
  h->prepare_capture();
  // The legacy default stream (what a caller gets from torch.cuda.current_stream() unless it made one) cannot be
  // captured: the loop then runs on a stream of the handle's own, ordered after the caller's stream by an event and
  // synchronised before this function returns.
  struct StreamSwap {
    gpb_handle* h;
    cudaStream_t user;
    bool on;
    ~StreamSwap() { if (on) h->s0 = user; }
  } swap{h, h->s0, h->s0 == nullptr || h->s0 == cudaStreamLegacy || h->s0 == cudaStreamPerThread};
  if (swap.on) {
    if (!h->s_loop) GPB_CUDA(cudaStreamCreateWithFlags(&h->s_loop, cudaStreamNonBlocking));
    cudaEvent_t e = h->next_event();
    GPB_CUDA(cudaEventRecord(e, swap.user));
    GPB_CUDA(cudaStreamWaitEvent(h->s_loop, e, 0));
    h->s0 = h->s_loop;
  }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  LoopResult res{0, 0, 0, 0};
  const int64_t l0 = h->launches;
  try {
    GPB_CUDA(cudaGraphCreate(&graph, 0));
    cudaGraphConditionalHandle cond;
    GPB_CUDA(cudaGraphConditionalHandleCreate(&cond, graph, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams np{};
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = cond;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    cudaGraphNode_t node;
    GPB_CUDA(cudaGraphAddNode(&node, graph, nullptr, 0, &np));
    cudaGraph_t body = np.conditional.phGraph_out[0];
    const auto t_cap0 = std::chrono::steady_clock::now();
    GPB_CUDA(cudaStreamBeginCaptureToGraph(h->s0, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    h->capturing = true;
    g_capturing = getenv("GPB_NO_PRIO") ? 0 : 1;
    try {
      iteration(cond);
    } catch (...) {
      h->capturing = false;
      g_capturing = 0;
      cudaGraph_t dummy = nullptr;
      cudaStreamEndCapture(h->s0, &dummy);
      throw;
    }
    h->capturing = false;
    g_capturing = 0;
    GPB_CUDA(cudaStreamEndCapture(h->s0, nullptr));
    res.launches_per_iteration = h->launches - l0;
    h->launches = l0;
    const auto t_cap1 = std::chrono::steady_clock::now();
    // per-node priorities (the look-ahead panel stream is a high-priority stream) only count with this flag
    GPB_CUDA(cudaGraphInstantiate(&exec, graph, getenv("GPB_NO_PRIO") ? 0 : cudaGraphInstantiateFlagUseNodePriority));
    const auto t_cap2 = std::chrono::steady_clock::now();
    GPB_CUDA(cudaGraphLaunch(exec, h->s0));
    if (getenv("GPB_DEBUG_LOOP"))
      fprintf(stderr, "[gpb] device loop: %lld launches per iteration, capture %.3f ms, instantiate %.3f ms, pdl %d\n",
              static_cast<long long>(res.launches_per_iteration), std::chrono::duration<double, std::milli>(t_cap1 - t_cap0).count(),
              std::chrono::duration<double, std::milli>(t_cap2 - t_cap1).count(), g_pdl);
    GPB_CUDA(cudaMemcpyAsync(hc_host, ctl_dev, sizeof(LoopCtl), cudaMemcpyDeviceToHost, h->s0));
    GPB_CUDA(cudaStreamSynchronize(h->s0));
    res.it = hc_host->it; res.status = hc_host->status; res.pivot = hc_host->pivot;
    if (res.it > 0 && trace_host) {
      GPB_CUDA(cudaMemcpyAsync(trace_host, trace_dev, static_cast<size_t>(res.it) * 16, cudaMemcpyDeviceToHost, h->s0));
      GPB_CUDA(cudaStreamSynchronize(h->s0));
    }
    h->launches += res.launches_per_iteration * res.it;
  } catch (...) {
    h->launches = l0;
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  cudaGraphExecDestroy(exec);
  cudaGraphDestroy(graph);
  return res;
}
// Diagnostic only (GPB_HOST_LOOP=1): the same iteration launched from the host with one synchronisation per iteration,
// i.e. the structure of the reference (GPpref.py:140-155) and of round 1.  Used to profile the kernels of an iteration
// with ncu (which does not open the body of a WHILE node) and to measure what the device loop saves.
template <class F>
LoopResult run_host_loop(gpb_handle* h, LoopCtl* ctl_dev, double* trace_dev, double* trace_host, double delta_f, int max_iter,
                         F& iteration) {
  LoopCtl* hc_host = reinterpret_cast<LoopCtl*>(h->pinned(sizeof(LoopCtl) + 64));
  *hc_host = LoopCtl{0, 0, 0, max_iter, delta_f};
  GPB_CUDA(cudaMemcpyAsync(ctl_dev, hc_host, sizeof(LoopCtl), cudaMemcpyHostToDevice, h->s0));
  LoopResult res{0, 0, 0, 0};
  for (;;) {
    const int64_t l0 = h->launches;
    iteration(0);
    res.launches_per_iteration = h->launches - l0;
    GPB_CUDA(cudaMemcpyAsync(hc_host, ctl_dev, sizeof(LoopCtl), cudaMemcpyDeviceToHost, h->s0));
    GPB_CUDA(cudaStreamSynchronize(h->s0));
    if (hc_host->status != 0) break;
  }
  res.it = hc_host->it; res.status = hc_host->status == 2 ? 2 : 0; res.pivot = hc_host->pivot;
  if (res.it > 0 && trace_host) GPB_CUDA(cudaMemcpy(trace_host, trace_dev, static_cast<size_t>(res.it) * 16, cudaMemcpyDeviceToHost));
  return res;
}
template <class F>
LoopResult run_device_loop(gpb_handle* h, LoopCtl* ctl_dev, double* trace_dev, double* trace_host, double delta_f, int max_iter,
                           F&& iteration) {
  if (getenv("GPB_HOST_LOOP")) return run_host_loop(h, ctl_dev, trace_dev, trace_host, delta_f, max_iter, iteration);
  try {
    return run_device_loop_once(h, ctl_dev, trace_dev, trace_host, delta_f, max_iter, iteration);
  } catch (const Error&) {
    // Programmatic dependent launches are the one construct of an iteration a graph body may refuse (no kernel has
    // run yet: capture or instantiation failed).  Same graph, plain dependencies.
    if (getenv("GPB_DEBUG_LOOP")) fprintf(stderr, "[gpb] device loop: first attempt failed (%s), retrying without PDL\n", h->err.c_str());
    if (g_pdl == 0) throw;
    cudaGetLastError();
    const int saved = g_pdl;
    g_pdl = 0;
    try {
      LoopResult r = run_device_loop_once(h, ctl_dev, trace_dev, trace_host, delta_f, max_iter, iteration);
      g_pdl = saved;
      return r;
    } catch (...) {
      g_pdl = saved;
      throw;
    }
  }
}

}  // namespace

#define LAP_BEGIN                                      \
  if (!h) return -1;                                   \
  try {                                                \
    GPB_CUDA(cudaSetDevice(h->device));                \
    ++h->ws_epoch;
#define LAP_END                                        \
  }                                                    \
  catch (const gpb::Error& e) {                        \
    h->err = e.msg;                                    \
    return -2;                                         \
  }                                                    \
  catch (const std::exception& e) {                    \
    h->err = e.what();                                 \
    return -3;                                         \
  }                                                    \
  return 0;

extern "C" {

int gpb_pref_derivatives(gpb_handle* h, const int64_t* uvi, const double* y, int64_t P, int64_t n, const double* f,
                         double sigma, int32_t grad_mode, double* W_out, double* g_out) {
  LAP_BEGIN
  GPB_REQUIRE(uvi && y && f && P > 0 && n > 0 && W_out && g_out, "null argument");
  PrefGraphHost g = build_pref_graph(uvi, P, n);
  PrefDev pd = upload_pref(h, g, uvi, y, P, n);
  h->aux0.ensure(static_cast<size_t>(n) * 16);
  h->aux2.ensure(static_cast<size_t>(n) * n * 8);
  double* fd = h->aux0.as<double>();
  double* gd = fd + n;
  GPB_CUDA(cudaMemcpyAsync(fd, f, n * 8, cudaMemcpyHostToDevice, h->s0));
  GPB_CUDA(cudaMemsetAsync(h->aux2.p, 0, static_cast<size_t>(n) * n * 8, h->s0));
  const double isq = 1.0 / (sigma * std::sqrt(2.0)), i2v = isq * isq;      // GPpref.py:53-54
  pref_pair_kernel<<<static_cast<unsigned>((P + 255) / 256), 256, 0, h->s0>>>(pd.uvi, pd.y, P, fd, isq, i2v, pd.dk, pd.wk);
  pref_row_kernel<<<static_cast<unsigned>((n + 63) / 64), 64, 0, h->s0>>>(n, pd.gr, pd.dk, pd.wk, fd, grad_mode,
                                                                            nullptr, 0, nullptr, gd, h->aux2.as<double>());
  GPB_CUDA(cudaGetLastError());
  h->launches += 2;
  GPB_CUDA(cudaMemcpyAsync(W_out, h->aux2.p, static_cast<size_t>(n) * n * 8, cudaMemcpyDeviceToHost, h->s0));
  GPB_CUDA(cudaMemcpyAsync(g_out, gd, n * 8, cudaMemcpyDeviceToHost, h->s0));
  GPB_CUDA(cudaStreamSynchronize(h->s0));
  LAP_END
}

int gpb_pref_log_marginal(gpb_handle* h, const int64_t* uvi, const double* y, int64_t P, int64_t n, const double* f,
                          const double* iK, double logdetK, double sigma, double* out) {
  LAP_BEGIN
  GPB_REQUIRE(uvi && y && f && iK && out && P > 0 && n > 0, "null argument");
  for (int64_t k = 0; k < 2 * P; ++k) GPB_REQUIRE(uvi[k] >= 0 && uvi[k] < n, "uvi index out of range");
  const int64_t ne = round_up(n, 2);                       // even pitch for the 16-byte loads of row_dot
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~size_t(255); return o; };
  const size_t o_uvi = take(P * 16), o_y = take(P * 8), o_f = take(ne * 8), o_f2 = take(ne * 8), o_t = take(ne * 8),
               o_sc = take(64 + 3 * PREF_PARTS * 8);
  h->aux1.ensure(off);
  h->aux2.ensure(static_cast<size_t>(n) * ne * 8);
  char* base = h->aux1.as<char>();
  double* fd = reinterpret_cast<double*>(base + o_f);
  double* f2 = reinterpret_cast<double*>(base + o_f2);
  double* td = reinterpret_cast<double*>(base + o_t);
  double* sc = reinterpret_cast<double*>(base + o_sc);
  GPB_CUDA(cudaMemsetAsync(base, 0, off, h->s0));
  GPB_CUDA(cudaMemsetAsync(h->aux2.p, 0, static_cast<size_t>(n) * ne * 8, h->s0));
  GPB_CUDA(cudaMemcpyAsync(base + o_uvi, uvi, P * 16, cudaMemcpyHostToDevice, h->s0));
  GPB_CUDA(cudaMemcpyAsync(base + o_y, y, P * 8, cudaMemcpyHostToDevice, h->s0));
  GPB_CUDA(cudaMemcpyAsync(fd, f, n * 8, cudaMemcpyHostToDevice, h->s0));
  GPB_CUDA(cudaMemcpyAsync(f2, f, n * 8, cudaMemcpyHostToDevice, h->s0));
  GPB_CUDA(cudaMemcpy2DAsync(h->aux2.p, ne * 8, iK, n * 8, n * 8, n, cudaMemcpyHostToDevice, h->s0));
  GPB_CUDA(cudaMemcpyAsync(sc, &logdetK, 8, cudaMemcpyHostToDevice, h->s0));
  launch_row_dot(h->aux2.as<double>(), ne, 0, fd, 0, n, ne, 0, td, 0, 1, h->s0);            // t = iK f
  const double isq = 1.0 / (sigma * std::sqrt(2.0));
  pref_terms_kernel<<<PREF_PARTS, 256, 0, h->s0>>>(reinterpret_cast<const int64_t*>(base + o_uvi),
                                                   reinterpret_cast<const double*>(base + o_y), P, n, isq, fd, f2, td, sc + 8);
  pref_finish_kernel<<<1, 256, 0, h->s0>>>(n, fd, f2, sc + 8, sc, sc + 2);
  GPB_CUDA(cudaGetLastError());
  h->launches += 2;
  double* host = h->pinned(64);
  GPB_CUDA(cudaMemcpyAsync(host, sc + 2, 16, cudaMemcpyDeviceToHost, h->s0));
  GPB_CUDA(cudaStreamSynchronize(h->s0));
  *out = host[1];
  LAP_END
}

int gpb_pref_laplace(gpb_handle* h, const int64_t* uvi, const double* y, int64_t P, const double* khyp, double sigma,
                     double delta_f, int32_t max_iter, int32_t grad_mode, int32_t use_f0, double* f_inout, double* lml,
                     int32_t* iters, double* trace, double* jitter, int32_t* info) {
  LAP_BEGIN
  GPB_REQUIRE(h->n > 0, "no training inputs: call gpb_set_train first");
  GPB_REQUIRE(uvi && y && khyp && f_inout && lml && iters && P > 0 && max_iter > 0, "null argument");
  const int64_t n = h->n, np = h->n_pad;
  if (info) *info = 0;
  PrefGraphHost graph = build_pref_graph(uvi, P, n);
  PrefDev pd = upload_pref(h, graph, uvi, y, P, n);
  GPB_CUDA(cudaEventRecord(h->tev[0], h->s0));

  // ---- K + eps I, its factor, log-determinant and explicit inverse (GPpref.py:121-135) ----
  double eps = 1e-6;                                                       // GPpref.py:123
  FactorMat m = laplace_mat(h, np);
  const double *ell, *hyp2;
  for (;;) {
    upload_kernel_params(h, khyp, eps, &ell, &hyp2);
    GPB_CUDA(cudaMemsetAsync(m.info, 0, 4, h->s0));
    build_kernel_matrix(h, ell, hyp2, m.A, m.ld, 1);
    m.rows_total = np;
    chol_sweep(h, m, true);
    if (read_info(h, m) == 0) break;
    eps *= 10.0;                                                           // GPpref.py:134
    if (!(eps < 1e8)) { if (info) *info = 1; throw Error{"K + eps*I is not positive definite for any jitter"}; }
  }
  if (jitter) *jitter = eps;
  h->scal.ensure(2048);
  double* sc = h->scal.as<double>();                 // [0] logdetK (= sum log diag L, GPpref.py:131)  [2..3] per-iteration out
  sum_log_kernel<<<1, 512, 0, h->s0>>>(m.diag, np, sc);
  ++h->launches;
  m.rows_total = 2 * np + TILE;
  chol_inverse_lower(h, m);                                                // iK (GPpref.py:129)
  h->aux2.ensure(static_cast<size_t>(np) * np * 8);
  double* iK = h->aux2.as<double>();
  {
    const int64_t nt64 = np / 64;
    mirror_lower_kernel<<<static_cast<unsigned>(nt64 * nt64), 256, 0, h->s0>>>(m.A, m.ld, iK, np, nt64);
    GPB_CUDA(cudaGetLastError());
    ++h->launches;
  }
  GPB_CUDA(cudaEventRecord(h->tev[1], h->s0));

  // ---- Newton iterations (GPpref.py:138-155) ----
  h->aux0.ensure(static_cast<size_t>(np) * 8 * 4);
  double* f = h->aux0.as<double>();
  double* f_new = f + np;
  double* t = f_new + np;
  GPB_CUDA(cudaMemsetAsync(f, 0, static_cast<size_t>(np) * 8 * 4, h->s0));
  if (use_f0) GPB_CUDA(cudaMemcpyAsync(f, f_inout, n * 8, cudaMemcpyHostToDevice, h->s0));
  const double isq = 1.0 / (sigma * std::sqrt(2.0)), i2v = isq * isq;      // GPpref.py:53-54
  FactorMat g = m;
  g.rows_total = np + 1;
  double* brow = g.A + np * g.ld;
  h->trace.ensure(static_cast<size_t>(max_iter) * 16);
  double* trace_dev = h->trace.as<double>();
  LoopCtl* ctl = reinterpret_cast<LoopCtl*>(sc + 16);
  double* part = sc + 32;                            // 3 x PREF_PARTS partial sums of the finish stage
  std::vector<double> trace_local;
  if (!trace) trace_local.resize(static_cast<size_t>(max_iter) * 2);
  double* trace_host = trace ? trace : trace_local.data();
  const LoopResult lr = run_device_loop(h, ctl, trace_dev, trace_host, delta_f, max_iter, [&](cudaGraphConditionalHandle cond) {
    pref_pair_kernel<<<static_cast<unsigned>((P + 255) / 256), 256, 0, h->s0>>>(pd.uvi, pd.y, P, f, isq, i2v, pd.dk, pd.wk);
    scale_copy_lower_kernel<<<static_cast<unsigned>(np), 256, 0, h->s0>>>(iK, np, g.A, g.ld, nullptr, 0.0);
    GPB_CUDA(cudaMemsetAsync(brow, 0, np * 8, h->s0));
    pref_row_kernel<<<static_cast<unsigned>((n + 63) / 64), 64, 0, h->s0>>>(n, pd.gr, pd.dk, pd.wk, f, grad_mode, g.A,
                                                                              g.ld, brow, nullptr, nullptr);
    GPB_CUDA(cudaGetLastError());
    h->launches += 3;
    GPB_CUDA(cudaMemsetAsync(g.info, 0, 4, h->s0));
    chol_sweep(h, g, true);                                  // G = L L^T, appended row <- L^-1 (W f + grad)
    trsv_lt(h, g, brow, f_new);                              // f_new = G^-1 (W f + grad)   (GPpref.py:143)
    launch_row_dot(iK, np, 0, f_new, 0, np, np, 0, t, 0, 1, h->s0);
    pref_terms_kernel<<<PREF_PARTS, 256, 0, h->s0>>>(pd.uvi, pd.y, P, n, isq, f_new, f, t, part);
    pref_finish_kernel<<<1, 256, 0, h->s0>>>(n, f_new, f, part, sc, sc + 2, ctl, trace_dev, g.info, cond);
    GPB_CUDA(cudaGetLastError());
    h->launches += 3;
  });
  const int it = lr.it;
  if (lr.status == 2) { if (info) *info = lr.pivot; throw Error{"K^-1 + W is not positive definite"}; }
  const double last_lml = it > 0 ? trace_host[2 * (it - 1) + 1] : 0.0;
  double* host = h->pinned(64);
  GPB_CUDA(cudaEventRecord(h->tev[2], h->s0));
  GPB_CUDA(cudaMemcpyAsync(f_inout, f, n * 8, cudaMemcpyDeviceToHost, h->s0));
  GPB_CUDA(cudaStreamSynchronize(h->s0));
  *lml = last_lml;
  *iters = it;
  for (float& x : h->timings) x = 0.f;
  GPB_CUDA(cudaEventElapsedTime(&h->timings[0], h->tev[0], h->tev[1]));
  GPB_CUDA(cudaEventElapsedTime(&h->timings[1], h->tev[1], h->tev[2]));
  GPB_CUDA(cudaEventElapsedTime(&h->timings[4], h->tev[0], h->tev[2]));
  h->lap_n = 0;
  {
    GPB_CUDA(cudaMemcpy(host, sc, 8, cudaMemcpyDeviceToHost));
    if (h->pref_state) pref_state_free(h->pref_state);
    PrefState* ps = new PrefState;
    h->pref_state = ps;
    h->pref_state_free = pref_state_free;
    ps->n = n; ps->P = P; ps->sigma = sigma; ps->eps = eps; ps->lml_ref = last_lml; ps->half_logdet_k = host[0];
    ps->grad_mode = grad_mode; ps->khyp.assign(khyp, khyp + h->d + 1); ps->pd = pd;
    h->state_epoch = h->ws_epoch;
  }
  LAP_END
}

int gpb_gpc_laplace(gpb_handle* h, const double* y, const double* khyp, int32_t link, double delta_f, int32_t max_iter,
                    int32_t use_f0, double* f_inout, double* lml, int32_t* iters, double* trace, double* jitter,
                    int32_t* info) {
  LAP_BEGIN
  GPB_REQUIRE(h->n > 0, "no training inputs: call gpb_set_train first");
  GPB_REQUIRE(y && khyp && f_inout && lml && iters && max_iter > 0, "null argument");
  const int64_t n = h->n, np = h->n_pad;
  if (info) *info = 0;
  h->lap_n = 0;
  GPB_CUDA(cudaEventRecord(h->tev[0], h->s0));
  FactorMat m = laplace_mat(h, np);
  h->aux2.ensure(static_cast<size_t>(np) * np * 8);
  double* K = h->aux2.as<double>();                      // full symmetric K + eps I
  double eps = 1e-6;
  const double *ell, *hyp2;
  for (;;) {
    upload_kernel_params(h, khyp, eps, &ell, &hyp2);
    build_kernel_matrix(h, ell, hyp2, K, np, 0);
    scale_copy_lower_kernel<<<static_cast<unsigned>(np), 256, 0, h->s0>>>(K, np, m.A, m.ld, nullptr, 0.0);
    GPB_CUDA(cudaMemsetAsync(m.info, 0, 4, h->s0));
    m.rows_total = np;
    chol_sweep(h, m, true);                              // positive-definiteness check of K (jitter loop)
    ++h->launches;
    if (read_info(h, m) == 0) break;
    eps *= 10.0;
    if (!(eps < 1e8)) { if (info) *info = 1; throw Error{"K + eps*I is not positive definite for any jitter"}; }
  }
  if (jitter) *jitter = eps;
  // vectors: f, f_new, b, sW, kb/t, a, y, g
  h->aux0.ensure(static_cast<size_t>(np) * 8 * 8);
  double* f = h->aux0.as<double>();
  double *f_new = f + np, *b = f + 2 * np, *sW = f + 3 * np, *tv = f + 4 * np, *a = f + 5 * np, *yd = f + 6 * np, *gv = f + 7 * np;
  GPB_CUDA(cudaMemsetAsync(f, 0, static_cast<size_t>(np) * 8 * 8, h->s0));
  GPB_CUDA(cudaMemcpyAsync(yd, y, n * 8, cudaMemcpyHostToDevice, h->s0));
  if (use_f0) GPB_CUDA(cudaMemcpyAsync(f, f_inout, n * 8, cudaMemcpyHostToDevice, h->s0));
  h->scal.ensure(2048);
  double* sc = h->scal.as<double>();
  FactorMat g = m;
  g.rows_total = np + 1;
  double* rrow = g.A + np * g.ld;
  const unsigned vb = static_cast<unsigned>((np + 255) / 256);
  h->trace.ensure(static_cast<size_t>(max_iter) * 16);
  double* trace_dev = h->trace.as<double>();
  LoopCtl* ctl = reinterpret_cast<LoopCtl*>(sc + 16);
  std::vector<double> trace_local;
  if (!trace) trace_local.resize(static_cast<size_t>(max_iter) * 2);
  double* trace_host = trace ? trace : trace_local.data();
  const LoopResult lr = run_device_loop(h, ctl, trace_dev, trace_host, delta_f, max_iter, [&](cudaGraphConditionalHandle cond) {
    gpc_terms_kernel<<<vb, 256, 0, h->s0>>>(link, yd, f, n, np, sW, b, nullptr);                 // Alg 3.1 lines 4, 6
    scale_copy_lower_kernel<<<static_cast<unsigned>(np), 256, 0, h->s0>>>(K, np, g.A, g.ld, sW, 1.0);   // B = I + sW K sW
    launch_row_dot(K, np, 0, b, 0, np, np, 0, tv, 0, 1, h->s0);                                  // K b
    vec_mul_kernel<<<vb, 256, 0, h->s0>>>(rrow, sW, tv, np);                                     // sW K b
    GPB_CUDA(cudaGetLastError());
    GPB_CUDA(cudaMemsetAsync(g.info, 0, 4, h->s0));
    chol_sweep(h, g, true);                                                                      // L = chol(B), row <- L^-1 (sW K b)
    trsv_lt(h, g, rrow, tv);                                                                     // L^T \ ( L \ (sW K b) )
    gpc_a_kernel<<<vb, 256, 0, h->s0>>>(a, b, sW, tv, np);                                       // line 7
    launch_row_dot(K, np, 0, a, 0, np, np, 0, f_new, 0, 1, h->s0);                               // line 8: f = K a
    gpc_finish_kernel<<<1, 1024, 0, h->s0>>>(link, yd, n, a, f_new, f, sc + 2, ctl, trace_dev, g.info, cond);
    GPB_CUDA(cudaGetLastError());
    h->launches += 7;
  });
  const int it = lr.it;
  if (lr.status == 2) { if (info) *info = lr.pivot; throw Error{"I + sW K sW is not positive definite"}; }
  const double last_obj = it > 0 ? trace_host[2 * (it - 1) + 1] : 0.0;
  double* host = h->pinned(64);
  // approximate log marginal likelihood at the returned f (Alg 3.1 line 10): W, L re-evaluated at f
  gpc_terms_kernel<<<vb, 256, 0, h->s0>>>(link, yd, f, n, np, sW, b, gv);
  scale_copy_lower_kernel<<<static_cast<unsigned>(np), 256, 0, h->s0>>>(K, np, g.A, g.ld, sW, 1.0);
  GPB_CUDA(cudaMemsetAsync(g.info, 0, 4, h->s0));
  g.rows_total = np;
  chol_sweep(h, g, true);
  sum_log_kernel<<<1, 512, 0, h->s0>>>(g.diag, np, sc);
  GPB_CUDA(cudaGetLastError());
  h->launches += 3;
  GPB_CUDA(cudaEventRecord(h->tev[1], h->s0));
  GPB_CUDA(cudaMemcpyAsync(host, sc, 8, cudaMemcpyDeviceToHost, h->s0));
  GPB_CUDA(cudaMemcpyAsync(f_inout, f, n * 8, cudaMemcpyDeviceToHost, h->s0));
  GPB_CUDA(cudaStreamSynchronize(h->s0));
  *lml = last_obj - host[0];
  *iters = it;
  for (float& x : h->timings) x = 0.f;
  GPB_CUDA(cudaEventElapsedTime(&h->timings[4], h->tev[0], h->tev[1]));
  // state for gpb_gpc_predict: L and Dinv (in h->A / h->Dinv), sW, grad, kernel parameters
  h->lap_n = n;
  h->state_epoch = h->ws_epoch;
  h->lap_link = link;
  h->lap_khyp.assign(khyp, khyp + h->d + 1);
  LAP_END
}

int gpb_gpc_predict(gpb_handle* h, const double* Z, int64_t mz, double* mu, double* var, double* prob) {
  LAP_BEGIN
  GPB_REQUIRE(h->lap_n == h->n && h->n > 0 && h->state_epoch + 1 == h->ws_epoch,
              "no Laplace state on this handle: call gpb_gpc_laplace first (any other call that uses the work space "
              "invalidates it)");
  GPB_REQUIRE(Z && mu && var && prob && mz > 0, "null argument");
  const int64_t np = h->n_pad;
  const int nt = static_cast<int>(np / TILE);
  FactorMat m = laplace_mat(h, np);            // same buffers: L of B is still in rows [0, np)
  double* f = h->aux0.as<double>();
  const double *sW = f + 3 * np, *gv = f + 7 * np;
  const double *ell, *hyp2;
  upload_kernel_params(h, h->lap_khyp.data(), 0.0, &ell, &hyp2);
  const double sf2 = h->lap_khyp[h->d];
  double* rows = m.A + (np + TILE) * m.ld;     // test rows live where grad.cu keeps U
  const int64_t cap = np;                      // rows available there
  h->outv.ensure(static_cast<size_t>(cap) * 8 * 3);
  double* o = h->outv.as<double>();
  for (int64_t z0 = 0; z0 < mz; z0 += cap) {
    const int64_t mc = mz - z0 < cap ? mz - z0 : cap;
    const int64_t mp64 = round_up(mc, 64);
    h->Zd.ensure(static_cast<size_t>(mc) * h->d * 8);
    GPB_CUDA(cudaMemcpyAsync(h->Zd.p, Z + z0 * h->d, static_cast<size_t>(mc) * h->d * 8, cudaMemcpyHostToDevice, h->s0));
    h->ZsT.ensure(static_cast<size_t>(h->d) * mp64 * 8);
    h->zsq.ensure(static_cast<size_t>(mp64) * 8);
    launch_se_prep(h->Zd.as<double>(), mc, h->d, ell, h->ZsT.as<double>(), mp64, h->zsq.as<double>(), 1, 0, 0, 0, h->s0);
    SeArgs a{};
    a.rT = h->ZsT.as<double>(); a.r_ld = mp64; a.r_sq = h->zsq.as<double>(); a.n_rows_valid = mc;
    a.cT = h->XsT.as<double>(); a.c_ld = np; a.c_sq = h->sq.as<double>(); a.n_cols_valid = h->n;
    a.d = h->d; a.out = rows; a.ld = m.ld; a.rows_pad = mp64; a.cols_pad = np;
    a.hyp_dev = hyp2; a.mode = 2; a.clip = 1;
    launch_se_build(a, 1, h->s0);                                                   // k*^T rows
    launch_row_dot(rows, m.ld, 0, gv, 0, mc, np, 0, o, 0, 1, h->s0);                // Alg 3.2 line 4: k*^T grad
    scale_cols_kernel<<<static_cast<unsigned>(mc), 256, 0, h->s0>>>(rows, m.ld, np, sW);
    GPB_CUDA(cudaGetLastError());
    h->launches += 4;
    m.rows_total = np + TILE + mc;
    SweepPlan plan;
    plan.factor = false;
    plan.extra_tile0 = nt + 1;
    plan.extra_tiles = static_cast<int>((mc + TILE - 1) / TILE);
    chol_sweep(h, m, plan);                                                         // rows <- (L^-1 sW k*)^T  (line 5)
    gpc_predict_finish_kernel<<<static_cast<unsigned>(mc), 256, 0, h->s0>>>(rows, m.ld, np, sf2, h->lap_link, o, o + cap, o + 2 * cap);
    GPB_CUDA(cudaGetLastError());
    ++h->launches;
    GPB_CUDA(cudaMemcpyAsync(mu + z0, o, mc * 8, cudaMemcpyDeviceToHost, h->s0));
    GPB_CUDA(cudaMemcpyAsync(var + z0, o + cap, mc * 8, cudaMemcpyDeviceToHost, h->s0));
    GPB_CUDA(cudaMemcpyAsync(prob + z0, o + 2 * cap, mc * 8, cudaMemcpyDeviceToHost, h->s0));
    GPB_CUDA(cudaStreamSynchronize(h->s0));
  }
  h->state_epoch = h->ws_epoch;                // the state is still current: predict again without refitting
  LAP_END
}

int gpb_pref_evidence(gpb_handle* h, double* evidence) {
  LAP_BEGIN
  GPB_REQUIRE(evidence, "null argument");
  PrefState* ps = pref_state_checked(h);
  pref_factor_at_mode(h, ps);
  // lml_ref = sum log Phi - f'K^-1 f/2 - (1/2) sum log diag L_K - n/2 log 2pi   (GPpref.py:90-94,131)
  // evidence = sum log Phi - f'K^-1 f/2 - sum log diag L_K - sum log diag chol(K^-1 + W)        (R&W eq. 3.32)
  *evidence = ps->lml_ref - 0.5 * ps->half_logdet_k + 0.5 * static_cast<double>(ps->n) * 1.8378770664093453 - ps->sum_log_g;
  h->state_epoch = h->ws_epoch;
  LAP_END
}

int gpb_pref_predict(gpb_handle* h, const double* Z, const double* Zb, int64_t mz, double* mean, double* var, double* prob) {
  LAP_BEGIN
  GPB_REQUIRE(Z && mean && var && mz > 0 && (prob || !Zb), "null argument");
  PrefState* ps = pref_state_checked(h);
  pref_factor_at_mode(h, ps);
  const int64_t np = h->n_pad;
  const int d = h->d;
  FactorMat m = laplace_mat(h, np);
  const double* iK = h->aux2.as<double>();
  double* f = h->aux0.as<double>();
  double* alpha = f + 2 * np;                                     // K^-1 f_hat
  launch_row_dot(iK, np, 0, f, 0, np, np, 0, alpha, 0, 1, h->s0);
  ++h->launches;
  const double *ell, *hyp2;
  upload_kernel_params(h, ps->khyp.data(), ps->eps, &ell, &hyp2);
  const double sf2 = ps->khyp[d];
  double* S1 = m.A + (np + TILE) * m.ld;                          // R = k* rows (np rows available)
  double* S2 = m.A + (2 * np + TILE) * m.ld;                      // T = R K^-1, then V = T L_G^-T
  TileMaps map_ik;
  make_tile_maps(&map_ik, iK, np, np, 1, np, np * np);
  const int64_t cap = np;
  h->outv.ensure(static_cast<size_t>(cap) * 8 * 5);
  double* o = h->outv.as<double>();                               // mean | var | prob | kss | q1
  h->Zd.ensure(static_cast<size_t>(cap) * d * 8 * 2);
  h->ZsT.ensure(static_cast<size_t>(d) * cap * 8 * 2);
  h->zsq.ensure(static_cast<size_t>(cap) * 8 * 2);
  for (int64_t z0 = 0; z0 < mz; z0 += cap) {
    const int64_t mc = mz - z0 < cap ? mz - z0 : cap;
    const int64_t mp64 = round_up(mc, 64);
    double* zd = h->Zd.as<double>();
    double* zsT = h->ZsT.as<double>();
    double* zsq = h->zsq.as<double>();
    auto cov_rows = [&](const double* Zsrc, int slot, double* dst) {
      GPB_CUDA(cudaMemcpyAsync(zd + slot * cap * d, Zsrc + z0 * d, static_cast<size_t>(mc) * d * 8, cudaMemcpyHostToDevice, h->s0));
      launch_se_prep(zd + slot * cap * d, mc, d, ell, zsT + slot * d * cap, mp64, zsq + slot * cap, 1, 0, 0, 0, h->s0);
      SeArgs a{};
      a.rT = zsT + slot * d * cap; a.r_ld = mp64; a.r_sq = zsq + slot * cap; a.n_rows_valid = mc;
      a.cT = h->XsT.as<double>(); a.c_ld = np; a.c_sq = h->sq.as<double>(); a.n_cols_valid = h->n;
      a.d = d; a.out = dst; a.ld = m.ld; a.rows_pad = mp64; a.cols_pad = np;
      a.hyp_dev = hyp2; a.mode = 2; a.clip = 1;                   // GPy RBF semantics, as in the fit (GPpref.py:122)
      launch_se_build(a, 1, h->s0);
      h->launches += 2;
    };
    cov_rows(Z, 0, S1);
    if (Zb) {
      cov_rows(Zb, 1, S2);
      rows_sub_kernel<<<static_cast<unsigned>(mc), 256, 0, h->s0>>>(S1, S2, m.ld, np);      // rows of f(b) - f(a)
      ++h->launches;
    }
    pref_kss_kernel<<<static_cast<unsigned>((mc + 255) / 256), 256, 0, h->s0>>>(zsT, Zb ? zsT + d * cap : nullptr, mp64, d, mc,
                                                                                sf2, o + 3 * cap);
    launch_row_dot(S1, m.ld, 0, alpha, 0, mc, np, 0, o, 0, 1, h->s0);                        // mean = R K^-1 f
    {
      GemmArgs a{};                                               // T = R * iK^T (iK symmetric), 64-tiles
      a.C = S2; a.ldc = m.ld; a.c_batch_stride = 0; a.rows_total = static_cast<int>(mp64);
      a.j0 = 0; a.j1 = static_cast<int>(np / 64); a.R = static_cast<int>(mp64 / 64); a.tri = 0; a.i0 = 0;
      a.ka0 = 0; a.kb0 = 0; a.nk = static_cast<int>(np / GEMM_KB);
      a.a_row0 = static_cast<int>(np + TILE); a.b_row0 = 0; a.epi = 0;
      launch_dmma_gemm(m.mapA.m64, map_ik.m64, a, 1, h->s0, 64);
    }
    rows_dot2_kernel<<<static_cast<unsigned>(mc), 256, 0, h->s0>>>(S1, S2, m.ld, np, o + 4 * cap);   // q1 = R K^-1 R'
    GPB_CUDA(cudaGetLastError());
    h->launches += 4;
    m.rows_total = 2 * np + TILE + mc;
    SweepPlan plan;
    plan.factor = false;
    plan.extra_tile0 = static_cast<int>((2 * np + TILE) / TILE);
    plan.extra_tiles = static_cast<int>((mc + TILE - 1) / TILE);
    chol_sweep(h, m, plan);                                       // V = T L_G^-T
    pref_predict_finish_kernel<<<static_cast<unsigned>(mc), 256, 0, h->s0>>>(S2, m.ld, np, o + 3 * cap, o + 4 * cap, o,
                                                                            2.0 * ps->sigma * ps->sigma, o + cap,
                                                                            Zb ? o + 2 * cap : nullptr);
    GPB_CUDA(cudaGetLastError());
    ++h->launches;
    GPB_CUDA(cudaMemcpyAsync(mean + z0, o, mc * 8, cudaMemcpyDeviceToHost, h->s0));
    GPB_CUDA(cudaMemcpyAsync(var + z0, o + cap, mc * 8, cudaMemcpyDeviceToHost, h->s0));
    if (Zb) GPB_CUDA(cudaMemcpyAsync(prob + z0, o + 2 * cap, mc * 8, cudaMemcpyDeviceToHost, h->s0));
    GPB_CUDA(cudaStreamSynchronize(h->s0));
  }
  h->state_epoch = h->ws_epoch;
  LAP_END
}

}  // extern "C"
