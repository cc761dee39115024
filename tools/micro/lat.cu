// Dependent-issue latencies of the FP64 operations on the panel's critical path (one warp, one SM).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, long long* clk, double x0) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 1024; ++i) {
    if (OP == 0) x = fma(x, y, 1e-9);
    if (OP == 1) x = x * y;
    if (OP == 2) x = rsqrt(x) + 1.0;
    if (OP == 3) x = sqrt(x) + 1.0;
    if (OP == 4) x = 1.0 / x + 1.0;
    if (OP == 5) { __syncthreads(); x += 1.0; }
    if (OP == 6) { float f = rsqrtf((float)x); x = (double)f + 1.0; }
    if (OP == 7) x = x + y;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { clk[0] = t1 - t0; }
  out[threadIdx.x] = x;
}
int main() {
  double* out; long long* clk; cudaMalloc(&out, 4096 * 8); cudaMalloc(&clk, 8);
  const char* names[] = {"DFMA", "DMUL", "rsqrt(double)+DADD", "sqrt(double)+DADD", "1/x+DADD", "__syncthreads(256)+DADD", "F2F+rsqrtf+F2F+DADD", "DADD"};
  for (int op = 0; op < 8; ++op) {
    int threads = op == 5 ? 256 : 32;
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: k<0><<<1, threads>>>(out, clk, 1.5); break; case 1: k<1><<<1, threads>>>(out, clk, 1.5); break;
        case 2: k<2><<<1, threads>>>(out, clk, 1.5); break; case 3: k<3><<<1, threads>>>(out, clk, 1.5); break;
        case 4: k<4><<<1, threads>>>(out, clk, 1.5); break; case 5: k<5><<<1, threads>>>(out, clk, 1.5); break;
        case 6: k<6><<<1, threads>>>(out, clk, 1.5); break; case 7: k<7><<<1, threads>>>(out, clk, 1.5); break;
      }
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("%-28s %7.1f clk per iteration\n", names[op], c / 1024.0);
  }
  return 0;
}
