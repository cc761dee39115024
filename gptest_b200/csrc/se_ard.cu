// se_ard.cu - covariance assembly for the squared-exponential ARD kernel.
//
//   K[i][j] = sf2 * exp(-0.5 * r2(i,j)) (+ sn2 on the diagonal),
//   r2 = |x_i/l|^2 + |x_j/l|^2 - 2 (x_i/l).(x_j/l)        (the reference's expanded form)
//
// follows SquaredExponential.compute_Kxx_matrix / compute_Kxz_matrix (GPr.py:99-110) and
// squared_distance (GPr.py:4-13): same scaling by division, same association
// (A2 + B2) - 2AB, same exp(-0.5*.) * sf2 + sn2*eye.  One fused pass: no N x N temporaries
// (the reference makes four), distance and exp never leave registers.
//
// Layout: 64 x 64 output tile per CTA, 256 threads, 4 x 4 cells per thread interleaved by 16 so
// that the 16 lanes of a half warp store 128 contiguous bytes of a row.  The scaled points are
// staged d-major in shared memory (conflict-free column reads, broadcast row reads).  In the
// symmetric mode only tiles on or below the diagonal are computed; the mirror tile is written
// through a padded shared-memory transpose so both writes are full-line coalesced.
//
// The kernel is a pure write stream (2 GiB at N=16384; a plain fill reaches 7.4 TB/s on this part), so
// what has to be kept small is the instruction count per element: ncu on the first version showed 179
// thread instructions per computed element (55 FP64 - 37 in the library exp -, 23 IMAD, 15 FSEL).
// Hence: table-driven exp (gpb_exp.cuh, 11 FP64), mode / clip as template parameters, row pointers
// hoisted, padding selects only in edge tiles.
#include "gpb_exp.cuh"
#include "gpb_kernels.cuh"

namespace gpb {

constexpr int ST = 64;       // output tile edge
constexpr int SDC = 8;       // dimensions staged per chunk

__global__ void __launch_bounds__(256) se_prep_kernel(const double* __restrict__ X, int64_t n, int d,
                                                      const double* __restrict__ ell, double* __restrict__ XsT,
                                                      int64_t ld_t, double* __restrict__ sq,
                                                      int64_t ell_bs, int64_t xs_bs, int64_t sq_bs) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int b = blockIdx.y;
  if (i >= ld_t) return;
  const double* l = ell + b * ell_bs;
  double* xt = XsT + b * xs_bs;
  double s = 0.0;
  for (int k = 0; k < d; ++k) {
    const double xs = (i < n) ? X[i * d + k] / l[k] : 0.0;     // GPr.py:100 scaledX = x / M
    xt[k * ld_t + i] = xs;
    s = fma(xs, xs, s);
  }
  sq[b * sq_bs + i] = s;
}

void launch_se_prep(const double* X, int64_t n, int d, const double* ell_dev, double* XsT,
                    int64_t ld_t, double* sq, int batch, int64_t ell_bs, int64_t xs_bs, int64_t sq_bs,
                    cudaStream_t st) {
  dim3 grid(static_cast<unsigned>((ld_t + 255) / 256), batch);
  se_prep_kernel<<<grid, 256, 0, st>>>(X, n, d, ell_dev, XsT, ld_t, sq, ell_bs, xs_bs, sq_bs);
  GPB_CUDA(cudaGetLastError());
}

// MODE 0: symmetric, both triangles  1: symmetric, lower tiles only  2: rectangular  3: rectangular, raw r^2
template <int MODE, int CLIP, int KIND>
__global__ void __launch_bounds__(256, 4) se_build_kernel(const SeArgs p) {
  __shared__ double xr[SDC][ST];
  __shared__ double xc[SDC][ST];
  __shared__ double tt[ST][ST + 1];
  __shared__ double etab[64];
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int b = blockIdx.y;
  exp_table_to_smem(etab);                      // made visible by the barriers of the staging loop below

  int64_t ti, tj;
  if (MODE >= 2) {
    const int64_t ntc = p.cols_pad / ST;
    ti = blockIdx.x / ntc;
    tj = blockIdx.x % ntc;
  } else {
    const double fi = (sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5;
    ti = static_cast<int64_t>(fi);
    while ((ti + 1) * (ti + 2) / 2 <= blockIdx.x) ++ti;
    while (ti * (ti + 1) / 2 > blockIdx.x) --ti;
    tj = blockIdx.x - ti * (ti + 1) / 2;
  }
  const int64_t i0 = ti * ST, j0 = tj * ST;
  const double* rT = p.rT + b * p.xs_batch_stride;
  const double* cT = p.cT + b * p.xs_batch_stride;

  double dot[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) dot[a][c] = 0.0;

  for (int d0 = 0; d0 < p.d; d0 += SDC) {
    const int dc = min(SDC, p.d - d0);
    __syncthreads();
    for (int e = t; e < dc * ST; e += 256) {
      const int k = e / ST, i = e % ST;
      xr[k][i] = rT[(d0 + k) * p.r_ld + i0 + i];
      xc[k][i] = cT[(d0 + k) * p.c_ld + j0 + i];
    }
    __syncthreads();
    for (int k = 0; k < dc; ++k) {
      double ra[4], ca[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) ra[a] = xr[k][ty + 16 * a];
#pragma unroll
      for (int c = 0; c < 4; ++c) ca[c] = xc[k][tx + 16 * c];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) dot[a][c] = fma(ra[a], ca[c], dot[a][c]);
    }
  }

  const double sf2 = p.hyp_dev[2 * b], sn2 = p.hyp_dev[2 * b + 1];
  const double* rsq = p.r_sq + b * p.sq_batch_stride;
  const double* csq = p.c_sq + b * p.sq_batch_stride;
  // -r2/2 = ab - (|a|^2/2 + |b|^2/2): the halvings are exact, so this is bit-identical to
  // -0.5 * ((A2 + B2) - 2 AB) of GPr.py:12,102 with one operation less per element
  double hr[4], hc[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) hr[a] = 0.5 * rsq[i0 + ty + 16 * a];
#pragma unroll
  for (int c = 0; c < 4; ++c) hc[c] = 0.5 * csq[j0 + tx + 16 * c];

  // one pointer per owned row, element (a, c) at rowp[a] + 16 c: no per-element address arithmetic
  double* out = p.out + b * p.out_batch_stride;
  double* rowp[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) rowp[a] = out + (i0 + ty + 16 * a) * p.ld + j0 + tx;
  const bool mirror = (MODE == 0) && (ti != tj);
  const bool interior = (i0 + ST <= p.n_rows_valid) && (j0 + ST <= p.n_cols_valid);
  const bool diag_tile = (MODE < 2) && (ti == tj);

  double val[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double x = dot[a][c] - (hr[a] + hc[c]);
      if (CLIP) x = fmin(x, 0.0);                                    // r2 clipped at 0 (GPy RBF semantics)
      val[a][c] = (MODE == 3) ? -2.0 * x : sf2 * radial<KIND>(x, etab);   // GPr.py:102 / :109 (mode 3: GPr.py:12 only)
    }
  if (diag_tile && ty == tx) {                                 // sn2 * eye: cells with a == c of the threads on the diagonal
#pragma unroll
    for (int a = 0; a < 4; ++a) val[a][a] += sn2;
  }
  if (!interior) {                                             // edge tiles only: identity (symmetric) / zero (rectangular) padding
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int64_t r = i0 + ty + 16 * a, cc = j0 + tx + 16 * c;
        if (r >= p.n_rows_valid || cc >= p.n_cols_valid) val[a][c] = (MODE < 2 && r == cc) ? 1.0 : 0.0;
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) rowp[a][16 * c] = val[a][c];
  if (mirror) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) tt[ty + 16 * a][tx + 16 * c] = val[a][c];
    __syncthreads();
    double* mp = out + (j0 + ty) * p.ld + i0 + tx;
    const int64_t rstep = 16 * p.ld;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
      for (int c = 0; c < 4; ++c) mp[16 * c] = tt[tx + 16 * c][ty + 16 * a];
      mp += rstep;
    }
  }
}

template <int MODE>
static void launch_mode(const SeArgs& a, dim3 grid, cudaStream_t st) {
  if (MODE < 3 && a.kind == 1) se_build_kernel<MODE, 1, 1><<<grid, 256, 0, st>>>(a);        // Matern: r^2 always clamped
  else if (MODE < 3 && a.kind == 2) se_build_kernel<MODE, 1, 2><<<grid, 256, 0, st>>>(a);
  else if (a.clip) se_build_kernel<MODE, 1, 0><<<grid, 256, 0, st>>>(a);
  else se_build_kernel<MODE, 0, 0><<<grid, 256, 0, st>>>(a);
}

void launch_se_build(const SeArgs& a, int batch, cudaStream_t st) {
  const int64_t tr = a.rows_pad / ST, tcn = a.cols_pad / ST;
  GPB_REQUIRE(a.rows_pad % ST == 0 && a.cols_pad % ST == 0, "se_build: padded extents must be multiples of 64");
  GPB_REQUIRE(a.kind >= 0 && a.kind <= 2, "se_build: unknown covariance kind");
  int64_t tiles = (a.mode >= 2) ? tr * tcn : tr * (tr + 1) / 2;
  if (tiles == 0) return;
  dim3 grid(static_cast<unsigned>(tiles), batch);
  switch (a.mode) {
    case 0: launch_mode<0>(a, grid, st); break;
    case 1: launch_mode<1>(a, grid, st); break;
    case 2: launch_mode<2>(a, grid, st); break;
    default: launch_mode<3>(a, grid, st); break;
  }
  GPB_CUDA(cudaGetLastError());
}

}  // namespace gpb
