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
#include "gpb_kernels.cuh"

namespace gpb {

constexpr int TP_THREADS = 256;
constexpr int TP_PITCH = TILE + 1;
constexpr int TP_SMEM = TILE * TP_PITCH * 8;

template <int JB>
__device__ __forceinline__ void potrf_block_steps(double (&c)[8][8], double* v, double* dsh, int* fail,
                                                  const int tr, const int tc) {
  for (int jj = 0; jj < 16; ++jj) {
    const int j = 16 * JB + jj;
    if (tr == jj && tc == jj) {
      const double ajj = c[JB][JB];
      if (!(ajj > 0.0)) atomicMin(fail, j);       // also catches NaN
      const double d = sqrt(ajj);
      c[JB][JB] = d;
      *dsh = d;
    }
    __syncthreads();
    const double invd = 1.0 / *dsh;
    if (tc == jj) {
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int r = tr + 16 * a;
        if (r != j) {
          c[a][JB] *= invd;
          v[r] = c[a][JB];
        } else {
          v[r] = invd;
        }
      }
    }
    __syncthreads();
    double vr[8], vc[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) vr[a] = v[tr + 16 * a];
#pragma unroll
    for (int b = JB; b < 8; ++b) vc[b] = v[tc + 16 * b];
#pragma unroll
    for (int b = JB; b < 8; ++b) {
      // column c = tc + 16 b is updated iff c > j
      const bool col_on = (b > JB) || (tc > jj);
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        // row r = tr + 16 a takes part iff r <= j (inverse rows) or r >= c (factor rows)
        const bool inv_row = (a < JB) || (a == JB && tr <= jj);
        const bool fac_row = (a > b) || (a == b && tr >= tc);
        if (col_on && (inv_row || fac_row)) c[a][b] = fma(-vr[a], vc[b], c[a][b]);
      }
    }
  }
}

__global__ void __launch_bounds__(TP_THREADS, 1) tile_potrf_inv_kernel(const TilePotrfArgs p) {
  extern __shared__ double S[];                 // [128][129] staging for the outputs
  __shared__ double v[TILE];
  __shared__ double dsh;
  __shared__ int fail;
  const int t = threadIdx.x;
  const int tc = t & 15, tr = t >> 4;
  const int batch = blockIdx.x;
  double* Ab = p.A + batch * p.a_batch_stride + static_cast<int64_t>(p.k) * TILE * p.lda + p.k * TILE;

  if (t == 0) fail = TILE;
  double c[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int r = tr + 16 * a, cc = tc + 16 * b;
      c[a][b] = (r >= cc) ? Ab[static_cast<int64_t>(r) * p.lda + cc] : 0.0;
    }
  __syncthreads();

  potrf_block_steps<0>(c, v, &dsh, &fail, tr, tc);
  potrf_block_steps<1>(c, v, &dsh, &fail, tr, tc);
  potrf_block_steps<2>(c, v, &dsh, &fail, tr, tc);
  potrf_block_steps<3>(c, v, &dsh, &fail, tr, tc);
  potrf_block_steps<4>(c, v, &dsh, &fail, tr, tc);
  potrf_block_steps<5>(c, v, &dsh, &fail, tr, tc);
  potrf_block_steps<6>(c, v, &dsh, &fail, tr, tc);
  potrf_block_steps<7>(c, v, &dsh, &fail, tr, tc);

#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) S[(tr + 16 * a) * TP_PITCH + tc + 16 * b] = c[a][b];
  __syncthreads();

  double* Dk = p.Dinv + batch * p.d_batch_stride + static_cast<int64_t>(p.k) * TILE * TILE;
  for (int idx = t; idx < TILE * TILE; idx += TP_THREADS) {
    const int r = idx >> 7, cc = idx & 127;
    if (cc <= r) Ab[static_cast<int64_t>(r) * p.lda + cc] = S[r * TP_PITCH + cc];
    // W[r][cc] = (L^-T)[cc][r]: strict upper cell S[cc][r] for cc < r, 1/L_rr on the diagonal
    double w = 0.0;
    if (cc < r) w = S[cc * TP_PITCH + r];
    else if (cc == r) w = 1.0 / S[r * TP_PITCH + r];
    Dk[idx] = w;
  }
  if (t < TILE) p.diag[batch * p.diag_batch_stride + p.k * TILE + t] = S[t * TP_PITCH + t];
  if (t == 0 && fail < TILE) atomicCAS(p.info + batch, 0, p.k * TILE + fail + 1);
}

void tile_potrf_init() {
  GPB_CUDA(cudaFuncSetAttribute(tile_potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM));
}

void launch_tile_potrf_inv(TilePotrfArgs a, int batch, cudaStream_t st) {
  tile_potrf_inv_kernel<<<batch, TP_THREADS, TP_SMEM, st>>>(a);
  GPB_CUDA(cudaGetLastError());
}

}  // namespace gpb
