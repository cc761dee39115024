// Latency / single-warp throughput of the instructions on the diagonal-tile kernel's critical path (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/lat2 tools/micro/lat2.cu
// Each test is a fully unrolled sequence timed with clock64() inside one warp of one CTA.
#include <cstdio>
#include <cuda_runtime.h>
#define N 64
template <int OP>
__global__ void k(double* out, long long* clk, double x0, int idx) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = x0 + i;
  __shared__ double sm[64];
  sm[threadIdx.x] = x;
  __syncwarp();
  long long t0, t1;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) : "d"(x), "d"(acc[0]), "d"(acc[15]) : "memory");
  if (OP == 0) {        // dependent DFMA
#pragma unroll
    for (int i = 0; i < N; ++i) x = fma(x, y, 1e-9);
  } else if (OP == 1) { // dependent DMUL
#pragma unroll
    for (int i = 0; i < N; ++i) x = x * y;
  } else if (OP == 2) { // 16 independent DFMA chains: single-warp issue rate
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i & 15] = fma(acc[i & 15], y, 1e-9);
  } else if (OP == 3) { // dependent 64-bit shuffle
#pragma unroll
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (idx + i) & 31);
  } else if (OP == 4) { // rsqrt.approx.ftz.f64 (MUFU.RSQ64H) dependent
#pragma unroll
    for (int i = 0; i < N; ++i) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  } else if (OP == 5) { // library rsqrt dependent
#pragma unroll
    for (int i = 0; i < N; ++i) x = rsqrt(x);
  } else if (OP == 6) { // STS -> syncwarp -> LDS round trip, dependent
#pragma unroll
    for (int i = 0; i < N; ++i) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + idx) & 31]; __syncwarp(); }
  } else if (OP == 7) { // 4 independent DFMA chains
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i & 3] = fma(acc[i & 3], y, 1e-9);
  } else if (OP == 8) { // dependent DMMA (accumulator chain)
    double c0 = x, c1 = y;
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(y), "d"(y));
    x = c0 + c1;
  } else if (OP == 9) { // 4 independent DMMA accumulators
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = x;
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i & 3][0]), "+d"(c[i & 3][1]) : "d"(y), "d"(y));
    x = c[0][0] + c[1][0] + c[2][1] + c[3][1];
  } else if (OP == 10) { // FSEL pair + DMUL dependent (select on double)
#pragma unroll
    for (int i = 0; i < N; ++i) { x = x * y; x = (threadIdx.x > (unsigned)(idx + i)) ? x : 1.0; }
  }
  double s = x;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) : "d"(s) : "memory");
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  out[threadIdx.x] = s;
}
int main() {
  double* out; long long* clk; cudaMalloc(&out, 4096 * 8); cudaMalloc(&clk, 8);
  const char* names[] = {"dependent DFMA", "dependent DMUL", "16 independent DFMA chains (issue interval)", "dependent SHFL.64", "MUFU.RSQ64H dependent",
                         "rsqrt(double) dependent", "STS+syncwarp+LDS+syncwarp dependent", "4 independent DFMA chains", "dependent DMMA", "4 independent DMMA accumulators",
                         "DMUL + select dependent"};
  for (int op = 0; op < 11; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: k<0><<<1, 32>>>(out, clk, 1.5, 1); break; case 1: k<1><<<1, 32>>>(out, clk, 1.5, 1); break;
        case 2: k<2><<<1, 32>>>(out, clk, 1.5, 1); break; case 3: k<3><<<1, 32>>>(out, clk, 1.5, 1); break;
        case 4: k<4><<<1, 32>>>(out, clk, 1.5, 1); break; case 5: k<5><<<1, 32>>>(out, clk, 1.5, 1); break;
        case 6: k<6><<<1, 32>>>(out, clk, 1.5, 1); break; case 7: k<7><<<1, 32>>>(out, clk, 1.5, 1); break;
        case 8: k<8><<<1, 32>>>(out, clk, 1.5, 1); break; case 9: k<9><<<1, 32>>>(out, clk, 1.5, 1); break;
        case 10: k<10><<<1, 32>>>(out, clk, 1.5, 1); break;
      }
      cudaDeviceSynchronize();
    }
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    printf("%-46s %7.1f clk per op\n", names[op], c / double(N));
  }
  return 0;
}
