// Probe of the sm_100a int8 tensor path used by csrc/ozaki.cu (tcgen05.mma kind::i8, operands in shared memory in the
// un-swizzled K-major canonical layout, accumulators in tensor memory):
//   1. correctness of the descriptor fields (which of LBO / SBO is the K-chunk stride), of one MMA with several
//      B slices concatenated along N, and of the full 7-slice x 7-slice "diagonal" pattern (10 MMAs per 32-deep K step);
//   2. issue rate of that pattern out of shared memory on all SMs (cycles per K step; 896 = the tensor-pipe floor).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I bayesian_cbf_b200/csrc -o build/umma_i8_probe tools/microbench/umma_i8_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "tc5.cuh"
using namespace tc5;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

constexpr int S = 7;            // slices per operand
constexpr int TM = 128, TN = 64;
constexpr int A_STEP = S * TM * 32;   // 28672 bytes per K step
constexpr int B_STEP = S * TN * 32;   // 14336
constexpr int B_PAD = 8192;           // slack behind the B slices (concatenated reads never leave the slices; safety)

__device__ __forceinline__ void issue_pattern(uint32_t tmem, uint32_t a_base, uint32_t b_base, uint32_t lbo, uint32_t sbo,
                                              bool first) {
  // slice a of A (128 x 32B) times slices b = 0..6-a of B, concatenated along N in groups of <= 4 (N <= 256);
  // product (a, b) lands in accumulator region a + b (64 TMEM columns each).
#pragma unroll
  for (int a = 0; a < S; ++a) {
    const uint64_t ad = smem_desc_kmajor(a_base + a * (TM * 32), lbo, sbo);
    int b = 0;
    while (b <= S - 1 - a) {
      int nb = S - a - b;
      if (nb > 4) nb = 4;
      const uint64_t bd = smem_desc_kmajor(b_base + b * (TN * 32), lbo, sbo);
      mma_s8(tmem + (a + b) * TN, ad, bd, idesc_s8(TM, TN * nb), (first && a == 0) ? 0u : 1u);
      b += nb;
    }
  }
}

// the same products with A in tensor memory: slice a is copied to TMEM columns 448 + 8 a right before its MMAs; copies
// and MMAs execute in issue order, so the copy of the NEXT K step's slice a may be issued as soon as this step's
// slice-a MMAs are
__device__ __forceinline__ void issue_pattern_ts(uint32_t tmem, uint32_t a_base, uint32_t b_base, bool first) {
#pragma unroll
  for (int a = 0; a < S; ++a) {
    tmem_cp_128x256b(tmem + 448 + 8 * a, smem_desc_kmajor(a_base + a * (TM * 32), 128, 256));
    int b = 0;
    while (b <= S - 1 - a) {
      int nb = S - a - b;
      if (nb > 4) nb = 4;
      const uint64_t bd = smem_desc_kmajor(b_base + b * (TN * 32), 128, 256);
      mma_s8_ts(tmem + (a + b) * TN, tmem + 448 + 8 * a, bd, idesc_s8(TM, TN * nb), (first && a == 0) ? 0u : 1u);
      b += nb;
    }
  }
}

// mode 0: A0 x B0 (N=64);  1: A0 x [B0..B3] (N=256);  2: full pattern over KS K steps
__global__ void __launch_bounds__(128, 1) probe_kernel(const int8_t* Ablob, const int8_t* Bblob, int KS, int mode,
                                                       uint32_t lbo, uint32_t sbo, int* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 2 * A_STEP;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, KS * (A_STEP + B_STEP));
    for (int ks = 0; ks < KS; ++ks) {
      bulk_g2s(sA + ks * A_STEP, Ablob + (size_t)ks * A_STEP, A_STEP, &bar_load);
      bulk_g2s(sB + ks * B_STEP, Bblob + (size_t)ks * B_STEP, B_STEP, &bar_load);
    }
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    if (mode == 0) {
      mma_s8(tmem, smem_desc_kmajor(smem_u32(sA), lbo, sbo), smem_desc_kmajor(smem_u32(sB), lbo, sbo), idesc_s8(TM, 64), 0);
    } else if (mode == 1) {
      mma_s8(tmem, smem_desc_kmajor(smem_u32(sA), lbo, sbo), smem_desc_kmajor(smem_u32(sB), lbo, sbo), idesc_s8(TM, 256), 0);
    } else if (mode == 2) {
      for (int ks = 0; ks < KS; ++ks)
        issue_pattern(tmem, smem_u32(sA + ks * A_STEP), smem_u32(sB + ks * B_STEP), lbo, sbo, ks == 0);
    } else {  // mode 3: A slices copied to tensor memory (columns 448..503), MMAs take A from there
      for (int ks = 0; ks < KS; ++ks) issue_pattern_ts(tmem, smem_u32(sA + ks * A_STEP), smem_u32(sB + ks * B_STEP), ks == 0);
    }
    mma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < 512; c += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 512 + c + j] = (int)r[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// pattern 0: the 10-MMA concatenated pattern; 1: 28 separate N=64 MMAs; 2: 7 MMAs of N=256 (same MACs as 28 x N=64)
__global__ void __launch_bounds__(128, 1) rate_kernel(int pattern, int iters, long long* cycles, int random_fill = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_mma;
  __shared__ uint64_t bar_ring[8];
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x / 32;
  for (int i = threadIdx.x; i < (A_STEP + B_STEP + B_PAD) / 4; i += blockDim.x) {
    uint32_t h = (i + 1) * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    ((uint32_t*)smem)[i] = random_fill ? h : 0x01010101u * (i & 3);
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar_mma, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&bar_ring[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base;
  if (threadIdx.x == 0) {
    const uint32_t a_base = smem_u32(smem), b_base = smem_u32(smem + A_STEP);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (pattern == 0) {
        issue_pattern(tmem, a_base, b_base, 128, 256, it == 0);
      } else if (pattern == 3) {  // as 0, plus what the real pipeline does per K step: fence, commit to a stage barrier
        tc_fence_after();
        issue_pattern(tmem, a_base, b_base, 128, 256, it == 0);
        mma_commit(&bar_ring[it & 7]);
      } else if (pattern == 4) {  // commit every second K step
        tc_fence_after();
        issue_pattern(tmem, a_base, b_base, 128, 256, it == 0);
        if (it & 1) mma_commit(&bar_ring[(it >> 1) & 7]);
      } else if (pattern == 5) {
        issue_pattern_ts(tmem, a_base, b_base, it == 0);
      } else if (pattern == 1) {
        for (int a = 0; a < S; ++a)
          for (int b = 0; b <= S - 1 - a; ++b)
            mma_s8(tmem + (a + b) * TN, smem_desc_kmajor(a_base + a * TM * 32, 128, 256),
                   smem_desc_kmajor(b_base + b * TN * 32, 128, 256), idesc_s8(TM, 64), (it == 0 && a == 0) ? 0u : 1u);
      } else {
        for (int a = 0; a < S; ++a)
          mma_s8(tmem + (a & 1) * 256, smem_desc_kmajor(a_base + a * TM * 32, 128, 256),
                 smem_desc_kmajor(b_base, 128, 256), idesc_s8(TM, 256), (it == 0 && a < 2) ? 0u : 1u);
      }
    }
    mma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

static size_t a_off(int ks, int s, int row, int k) {  // k in [0,32)
  return (size_t)ks * A_STEP + s * (TM * 32) + (row / 8) * 256 + (k / 16) * 128 + (row % 8) * 16 + (k % 16);
}
static size_t b_off(int ks, int s, int col, int k) {
  return (size_t)ks * B_STEP + s * (TN * 32) + (col / 8) * 256 + (k / 16) * 128 + (col % 8) * 16 + (k % 16);
}

int main() {
  const int KS = 2;
  std::vector<int8_t> A(S * TM * KS * 32), B(S * TN * KS * 32), Ablob(KS * A_STEP), Bblob(KS * B_STEP);
  srand(1);
  for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
  for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
  auto Aat = [&](int s, int r, int k) { return (int)A[(s * TM + r) * KS * 32 + k]; };
  auto Bat = [&](int s, int c, int k) { return (int)B[(s * TN + c) * KS * 32 + k]; };
  for (int ks = 0; ks < KS; ++ks)
    for (int s = 0; s < S; ++s) {
      for (int r = 0; r < TM; ++r)
        for (int k = 0; k < 32; ++k) Ablob[a_off(ks, s, r, k)] = (int8_t)Aat(s, r, ks * 32 + k);
      for (int c = 0; c < TN; ++c)
        for (int k = 0; k < 32; ++k) Bblob[b_off(ks, s, c, k)] = (int8_t)Bat(s, c, ks * 32 + k);
    }
  int8_t *dA, *dB;
  int* dOut;
  CK(cudaMalloc(&dA, Ablob.size()));
  CK(cudaMalloc(&dB, Bblob.size()));
  CK(cudaMalloc(&dOut, 128 * 512 * 4));
  CK(cudaMemcpy(dA, Ablob.data(), Ablob.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bblob.data(), Bblob.size(), cudaMemcpyHostToDevice));
  const int smem_bytes = 2 * A_STEP + 2 * B_STEP + B_PAD;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  std::vector<int> out(128 * 512);
  for (int variant = 0; variant < 2; ++variant) {
    const uint32_t lbo = variant == 0 ? 128 : 256, sbo = variant == 0 ? 256 : 128;
    for (int mode = 0; mode < 4; ++mode) {
      if (mode == 3 && variant == 1) continue;
      CK(cudaMemset(dOut, 0xff, 128 * 512 * 4));
      probe_kernel<<<1, 128, smem_bytes>>>(dA, dB, mode >= 2 ? KS : 1, mode, lbo, sbo, dOut);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d mode %d: kernel failed: %s\n", variant, mode, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
      long long bad = 0, total = 0;
      int first_bad_r = -1, first_bad_c = -1, got = 0, want = 0;
      auto check = [&](int r, int col, long long ref) {
        ++total;
        if (out[r * 512 + col] != (int)ref) {
          if (!bad) { first_bad_r = r; first_bad_c = col; got = out[r * 512 + col]; want = (int)ref; }
          ++bad;
        }
      };
      if (mode == 0 || mode == 1) {
        const int nb = mode == 0 ? 1 : 4;
        for (int r = 0; r < TM; ++r)
          for (int b = 0; b < nb; ++b)
            for (int c = 0; c < TN; ++c) {
              long long ref = 0;
              for (int k = 0; k < 32; ++k) ref += Aat(0, r, k) * Bat(b, c, k);
              check(r, b * TN + c, ref);
            }
      } else {
        for (int r = 0; r < TM; ++r)
          for (int d = 0; d < S; ++d)
            for (int c = 0; c < TN; ++c) {
              long long ref = 0;
              for (int a = 0; a <= d; ++a)
                for (int k = 0; k < KS * 32; ++k) ref += Aat(a, r, k) * Bat(d - a, c, k);
              check(r, d * TN + c, ref);
            }
      }
      printf("variant %d (lbo=%u sbo=%u) mode %d: %lld / %lld mismatches", variant, lbo, sbo, mode, bad, total);
      if (bad) printf("  first at row %d col %d: got %d want %d", first_bad_r, first_bad_c, got, want);
      printf("\n");
    }
  }
  // ---- issue rate
  long long* dCyc;
  CK(cudaMalloc(&dCyc, 148 * 8));
  std::vector<long long> cyc(148);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const char* names[6] = {"10-MMA concatenated pattern", "28 separate N=64 MMAs", "7 MMAs of N=256",
                          "10-MMA + fence + commit / step", "10-MMA + commit / 2 steps", "10-MMA, A via tcgen05.cp -> TMEM"};
  for (int grid : {1, 148}) {
    for (int pattern = 0; pattern < 6; ++pattern) {
      const int iters = 20000;
      rate_kernel<<<grid, 128, 200 * 1024>>>(pattern, 200, dCyc);   // warm-up
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      rate_kernel<<<grid, 128, 200 * 1024>>>(pattern, iters, dCyc);
      CK(cudaEventRecord(e1));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("rate pattern %d failed: %s\n", pattern, cudaGetErrorString(e)); return 1; }
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(cyc.data(), dCyc, grid * 8, cudaMemcpyDeviceToHost));
      long long mn = cyc[0], mx = cyc[0];
      for (int i = 0; i < grid; ++i) { mn = cyc[i] < mn ? cyc[i] : mn; mx = cyc[i] > mx ? cyc[i] : mx; }
      const double macs = 28.0 * 128 * 64 * 32 * iters * grid;
      printf("rate grid=%3d %-30s: %.1f .. %.1f cycles per K step (floor 896), %.3f ms, %.1f int8 TOPS\n", grid,
             names[pattern], (double)mn / iters, (double)mx / iters, ms, 2 * macs / (ms * 1e-3) / 1e12);
    }
  }
  // ---- sustained rate under the power cap (about 2 s per run): the ceiling of any kernel built on this pattern
  for (int random_fill = 0; random_fill < 2; ++random_fill)
    for (int pattern : {0, 2}) {
      const int iters = 3000000;
      CK(cudaEventRecord(e0));
      rate_kernel<<<148, 128, 200 * 1024>>>(pattern, iters, dCyc, random_fill);
      CK(cudaEventRecord(e1));
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("sustained pattern %d failed: %s\n", pattern, cudaGetErrorString(e)); return 1; }
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      CK(cudaMemcpy(cyc.data(), dCyc, 148 * 8, cudaMemcpyDeviceToHost));
      const double macs = 28.0 * 128 * 64 * 32 * (double)iters * 148;
      printf("sustained %-28s %s digits: %.1f cycles per K step, %.1f ms, %.1f int8 TOPS, mean SM clock %.0f MHz\n",
             names[pattern], random_fill ? "random" : "constant", (double)cyc[0] / iters, ms, 2 * macs / (ms * 1e-3) / 1e12,
             (double)cyc[0] / (ms * 1e-3) / 1e6);
    }
  return 0;
}
