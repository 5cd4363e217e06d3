// Shared device helpers for the MVGP hot path (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#ifndef __CUDA_ARCH_LIST__
#endif

namespace bcbf {

constexpr int kBlk = 128;  // block edge used by every blocked algorithm (Cholesky panels, Linv tiles)

// ---- error plumbing -----------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define BCBF_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t _e = (call);                                                   \
    if (_e != cudaSuccess) return ::bcbf::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

// every kernel launch of the library goes through one of these (bench.py reports the count as gpu_launches)
extern unsigned long long g_launch_count;
#define BCBF_LAUNCH_CHECK()        \
  do {                             \
    ++::bcbf::g_launch_count;      \
    BCBF_CUDA(cudaGetLastError()); \
  } while (0)

#define BCBF_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::bcbf::set_last_error(__VA_ARGS__);      \
      return BCBF_ERR_INVALID;                  \
    }                                           \
  } while (0)

// ---- FP64 tensor-core MMA: D(8x8) += A(8x4, row) * B(4x8, col) ------------------------------------
// Fragment ownership (lane = 0..31):  A[m = lane/4][k = lane%4],  B[k = lane%4][n = lane/4],
//                                     C[m = lane/4][n = 2*(lane%4) + {0,1}].
// On sm_100a every f64 mma.sync shape lowers to SASS DMMA.8x8x4, so this is the hardware primitive.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// ---- cp.async (LDGSTS) 16-byte copies with zero-fill predicate ------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
  int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem_src), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// G[i, q] = sum_t UH[i, t] B[t, q] as one FMA chain over t — the ONE definition (Gram kernels, the residual of the alpha
// refinement and the model handle's G rows must agree bit for bit)
__device__ __forceinline__ double g_entry(const double* __restrict__ uh_row, const double* __restrict__ Bm, int p, int q) {
  double g = 0.0;
  for (int t = 0; t < p; ++t) g = __fma_rn(uh_row[t], Bm[t * p + q], g);
  return g;
}

// ---- per-device scratch ordering ------------------------------------------------------------------------------------
// Several entry points (posterior_*, oz_*, potrf, trtri) use library-owned per-device scratch buffers and, for potrf, a set
// of look-ahead streams and events.  A ScratchScope at the top of such an entry point makes that safe for callers on
// different streams / host threads of one device: a per-device host mutex is held while the call enqueues its work, the
// caller's stream first waits (on the GPU) for the previous scratch user if that ran on another stream, and an event is
// recorded when the call has enqueued everything.  Nested entry points (potrf -> oz_update) join the outer scope.
struct ScratchScope {
  explicit ScratchScope(cudaStream_t s);
  ~ScratchScope();
  ScratchScope(const ScratchScope&) = delete;
  ScratchScope& operator=(const ScratchScope&) = delete;
  cudaStream_t stream;
  int dev;
};

inline int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

}  // namespace bcbf
