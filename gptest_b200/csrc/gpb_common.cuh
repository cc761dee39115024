// gpb_common.cuh - shared helpers for libgpb200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgpb200 is written for sm_100a (B200) only"
#endif

namespace gpb {

constexpr int TILE = 128;          // tile edge of the blocked factorisation (elements)
constexpr int GEMM_KB = 16;        // k-slab per pipeline stage: 16 doubles = one 128-byte TMA row
constexpr int GEMM_STAGES = 4;

// ---- device-side PTX wrappers ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 2-D / 3-D TMA tile loads (global -> shared, completion on an mbarrier). SASS: UTMALDG.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// FP64 tensor-core MMA: D(8x8) += A(8x4,row) * B(4x8,col).  SASS: DMMA.8x8x4.
// lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- host-side error plumbing -----------------------------------------------------------
struct Error {
  std::string msg;
};
#define GPB_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      throw gpb::Error{std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " +    \
                       __FILE__ + ":" + std::to_string(__LINE__)};                            \
    }                                                                                         \
  } while (0)
#define GPB_REQUIRE(cond, text)                                                               \
  do {                                                                                        \
    if (!(cond)) throw gpb::Error{std::string("invalid argument: ") + (text)};               \
  } while (0)

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// ---- programmatic dependent launch -------------------------------------------------------
// The panel of the factorisation and the substitution sweeps are chains of small dependent kernels
// (tile_potrf -> TRSM -> update -> tile_potrf ...).  Launched with the programmatic-serialisation attribute, the
// next kernel of the chain becomes resident while its predecessor drains: launch latency and the prologue
// (barrier initialisation, table staging) leave the critical path.  Every kernel launched this way calls
// pdl_trigger() first and pdl_wait() before its first access to global memory; both are no-ops for a plain launch.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

extern int g_pdl;       // 0: off  1: the small launches of dependent chains (default)  2: every launch (dmma_gemm.cu)
extern int g_capturing; // != 0 while a Newton iteration is being captured into a graph body (laplace.cu): launches then
                        // carry their stream's priority as an explicit attribute, so that the look-ahead panel keeps
                        // its precedence over the trailing update inside the graph

template <class... KArgs, class... Args>
inline void launch_chain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                         Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (g_capturing) {
    int prio = 0;
    if (cudaStreamGetPriority(st, &prio) == cudaSuccess) {
      at[na].id = cudaLaunchAttributePriority;
      at[na].val.priority = prio;
      ++na;
    }
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  GPB_CUDA(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
}

}  // namespace gpb
