// Phase clocks of the diagonal-tile kernel (tile_potrf3.cu) on one tile: build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o tools/micro/tp3_bench tools/micro/tp3_bench.cu
// and run on the GPU box.  Prints, per phase boundary, the clock64() of every warp relative to the kernel start.
#include <cmath>
#include <cstdio>
#include <vector>
#include "../../gptest_b200/csrc/tile_potrf3.cu"

int main(int argc, char** argv) {
  const int refine = argc > 1 ? atoi(argv[1]) : 1;
  const int n = 128;
  std::vector<double> A(n * n), M(n * n);
  srand(1);
  for (auto& x : M) x = rand() / double(RAND_MAX) - 0.5;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0;
      for (int k = 0; k < n; ++k) s += M[i * n + k] * M[j * n + k];
      A[i * n + j] = s / n + (i == j ? 1.0 : 0.0);
    }
  double *dA, *dW, *dd;
  int* dinfo;
  long long* ddbg;
  cudaMalloc(&dA, n * n * 8); cudaMalloc(&dW, n * n * 8); cudaMalloc(&dd, n * 8); cudaMalloc(&dinfo, 4); cudaMalloc(&ddbg, (24 * 8 + 32 + 96) * 8);
  gpb::tile_potrf3_init();
  gpb::TilePotrfArgs a{};
  a.A = dA; a.lda = n; a.a_batch_stride = 0; a.k = 0; a.Dinv = dW; a.d_batch_stride = 0; a.diag = dd; a.diag_batch_stride = 0;
  a.info = dinfo; a.pdl = 0; a.dbg = ddbg;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int it = 0; it < 5; ++it) {
    cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
    cudaMemset(dinfo, 0, 4);
    cudaMemset(ddbg, 0, (24 * 8 + 32 + 96) * 8);
    cudaEventRecord(e0);
    gpb::launch_tile_potrf3(a, 1, 0, false, refine != 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(err)); return 1; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  std::vector<double> L(n * n), W(n * n);
  std::vector<long long> dbg(24 * 8 + 32 + 96);
  cudaMemcpy(L.data(), dA, n * n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(W.data(), dW, n * n * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(dbg.data(), ddbg, (24 * 8 + 32 + 96) * 8, cudaMemcpyDeviceToHost);
  // check: L L^T = A (lower), W L = I
  double e1m = 0, e2m = 0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = 0, w = 0;
      for (int k = 0; k <= j; ++k) s += L[i * n + k] * L[j * n + k];
      for (int k = j; k <= i; ++k) w += W[i * n + k] * L[k * n + j];
      e1m = fmax(e1m, fabs(s - A[i * n + j]));
      e2m = fmax(e2m, fabs(w - (i == j ? 1.0 : 0.0)));
    }
  printf("refine %d  kernel %.2f us (event, best of 5)   |LL^T-A| %.2e  |WL-I| %.2e\n", refine, best * 1e3, e1m, e2m);
  const char* names[24] = {"start", "block0 loaded", "", "", "P0 own", "P0 bar", "T0", "S0", "P1 own", "P1 bar", "T1", "S1", "P2 own", "P2 bar",
                           "T2", "S2", "P3 own", "P3 bar", "", "", "", "", "tail", "end"};
  const long long t0 = dbg[0];
  for (int s = 0; s < 24; ++s) {
    if (!names[s][0]) continue;
    printf("%-14s", names[s]);
    for (int w = 0; w < 8; ++w) printf(" %7lld", dbg[s * 8 + w] ? dbg[s * 8 + w] - t0 : -1);
    printf("\n");
  }
  printf("factoring warp, block 1 (per panel: start, steps begin, steps end, stored; then end):\n");
  for (int s = 0; s < 17; ++s) printf(" %lld", dbg[24 * 8 + s] - dbg[24 * 8]);
  printf("\n");
  printf("workers (per phase kb=1..3: sub-step a done, barrier passed, sub-step b done), relative to kernel start:\n");
  for (int s = 0; s < 12; ++s) { if (s % 4 == 3) continue; printf("kb%d.%d", s / 4 + 1, s % 4); for (int w = 0; w < 8; ++w) printf(" %7lld", dbg[24 * 8 + 32 + s * 8 + w] ? dbg[24 * 8 + 32 + s * 8 + w] - t0 : -1); printf("\n"); }
  return 0;
}
