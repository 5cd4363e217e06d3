// Batched posterior of F(x) / F(x)u over many query states.
//
// The dominant cost is  S(x) = frakB(x)^T Kb^{-1} frakB(x) = V^T V,  V = Linv frakB(x),
// frakB(x)[i, q] = k(X_i, x) G[i, q]  — N^2 p flops per query, a dense FP64 contraction that runs on the
// DMMA (FP64 tensor) pipe.  post_var_kernel is a persistent fused kernel: one CTA owns a tile of TQ queries
// (BN = p*TQ columns), walks ALL 128-row blocks I of Linv and, for each, the k-range [0,(I+1)*128):
//     - A operand: Linv[I-block, k..k+16) via a 3-stage cp.async pipeline (L2-resident: every CTA streams
//       the same Linv tiles in the same order, so HBM sees Linv once per wave);
//     - B operand: built on the fly, never stored in HBM: Kstar[k, query] * G[k, q] written straight into the
//       fragment-friendly shared tile;
//     - accumulators (128 x BN) stay in registers; after each row block the per-query p x p Gram
//       sum_rows V[:,q]V[:,r] is reduced with warp shuffles into shared memory (deterministic order);
//     - V itself never touches HBM.
// Algorithmic work per query: N^2 p (+ N p lower order) flops; HBM bytes per query: 8 N (its Kstar column,
// re-read (nb+1)/2 times from L2/HBM) + outputs.
#include "../../include/bcbf.h"
#include "common.cuh"

#include <utility>
#include <vector>

namespace bcbf {

// k extent of one pipeline stage is PostCfg::BK (32 or 64); A tile row stride BK + 4 doubles (== 4 mod 16: conflict-free
// fragment loads)
constexpr int kPMmaWarps = 8;                   // consumer warps: 2 (rows) x 4 (query groups), 64 x (P*H*8) each
constexpr int kPProdWarps = 4;                  // producer warps, one per SM sub-partition (warp id % 4): each streams a
                                                // quarter of the L^-1 / K* tiles with cp.async so that no sub-partition's
                                                // DMMA warps are slowed more than the others
constexpr int kPThreads = 32 * (kPMmaWarps + kPProdWarps);

// P = columns per query; H = 8-query fragments per warp; GMUL: B operand = K*[k,t] * G[k,q] formed in registers
// (false: the K* operand is used as is — fold-in form, where kb* already carries the (G . uh) factor).
template <int P_, int H_, bool GMUL_>
struct PostCfg {
  static constexpr int P = P_;
  static constexpr int H = H_;
  static constexpr bool GMUL = GMUL_;
  static constexpr int QW = 8 * H;                   // queries per warp column
  static constexpr int TQ = 4 * QW;                  // queries per CTA
  static constexpr int NF = P * H;                   // n-fragments per warp
  // 32-query tiles (p = 3, 4) can afford 64-deep stages (2 x 88 KB): half as many stage hand-overs per flop
  static constexpr int BK = (TQ <= 32) ? 64 : 32;
  static constexpr int AStride = BK + 4;
  static constexpr int AElems = 128 * AStride;
  static constexpr int KStride = TQ + 4;             // == 4 (mod 16): conflict-free fragment reads
  static constexpr int KElems = BK * KStride;
  static constexpr int GElems = BK * 4;              // packed G rows of the stage (BK * p doubles, p <= 4)
  static constexpr int NPair = P * (P + 1) / 2;
  static constexpr int StageElems = AElems + KElems + GElems;
  static constexpr int Stages = (TQ <= 32) ? 2 : 3;
  static constexpr int SmemBytes = (Stages * StageElems + 2 * TQ * NPair) * (int)sizeof(double) + 64;
  static_assert(TQ % 16 == 0, "tile width must keep the padded stride at 4 mod 16");
  static_assert(SmemBytes <= 232448, "exceeds the 227 KB of shared memory a CTA may use");
};

struct PostArgs {
  const double* Linv; int ld; int Npad;
  const double* Kstar; int ldks;   // (Npad, ldks): K* (GMUL) or kb* (fold-in)
  const double* G;       // (Npad, P) row-major, pad rows zero (GMUL only)
  int Q;
  double* Spart;         // [nsplit][Qpad][NPair] partial Gram sums
  int Qpad; int nsplit;
  unsigned long long* dbg;  // optional pipeline counters (bcbf_debug_counters); NULL in production
};

// ---- mbarrier helpers (CTA-scope producer/consumer pipeline; no __syncthreads in the steady state) ------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
// arrival that fires when all cp.async issued so far by this thread have landed (counted in the init count: .noinc)
__device__ __forceinline__ void mbar_arrive_cp_async(unsigned long long* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n"
      "WAIT_LOOP:\n"
      " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      " @p bra WAIT_DONE;\n"
      " bra WAIT_LOOP;\n"
      "WAIT_DONE:\n}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// Warp-specialised persistent kernel.  Warps 0..7 issue DMMA only (plus NF DMULs per 8*NF DMMAs to form the frakB
// fragments); warp 8 streams every operand with cp.async.  Stage s of the ring is handed over with full[s]
// (32 cp.async-completion arrivals) and returned with empty[s] (one arrival per MMA warp).
template <class Cfg>
__global__ void __launch_bounds__(kPThreads, 1) post_var_kernel(PostArgs a) {
  constexpr int P = Cfg::P, QW = Cfg::QW, TQ = Cfg::TQ, H = Cfg::H, NF = Cfg::NF, KS = Cfg::KStride;
  constexpr int NPair = Cfg::NPair, S = Cfg::Stages, BK = Cfg::BK, AS = Cfg::AStride;
  extern __shared__ __align__(16) double smem[];
  double* Ssm = smem + S * Cfg::StageElems;  // [2][TQ][NPair]
  unsigned long long* full = reinterpret_cast<unsigned long long*>(Ssm + 2 * TQ * NPair);
  unsigned long long* empty = full + S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * TQ;
  const int nb = a.Npad / kBlk;
  const int split = blockIdx.y, nsplit = a.nsplit;
  constexpr int kStagesPerBlk = kBlk / BK;

  const long long tk0 = a.dbg ? clock64() : 0;
  for (int i = tid; i < 2 * TQ * NPair; i += kPThreads) Ssm[i] = 0.0;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full + s, 32 * kPProdWarps);
      mbar_init(empty + s, kPMmaWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // iteration space: row blocks I = split, split+nsplit, ... ; stages kt in [0, (I+1)*4)
  auto stA = [&](int s) { return smem + s * Cfg::StageElems; };
  auto stK = [&](int s) { return smem + s * Cfg::StageElems + Cfg::AElems; };
  auto stG = [&](int s) { return smem + s * Cfg::StageElems + Cfg::AElems + Cfg::KElems; };

  if (warp >= kPMmaWarps) {
    // ================= producers: L^-1[I rows, k..k+BK) -> As[128][BK+4]; K*[k..k+BK, q0..q0+TQ) -> Ks[BK][TQ+4];
    //                   G[k..k+32, :] -> Gs (packed).  16-byte cp.async (LDGSTS) only; completion of a lane's copies
    //                   arrives on full[s] (cp.async.mbarrier.arrive).  Measured alternative: one 256-byte
    //                   cp.async.bulk (TMA unit, UBLKCP) per tile row — 161 small bulk copies per stage made the producer
    //                   the bottleneck (25.6 vs 32.3 TFLOP/s), so the LDGSTS path is kept. ======
    const int pw = warp - kPMmaWarps;
    int slot = 0;
    unsigned phase = 0;
    const double* gK0 = a.Kstar + q0;
    for (int I = split; I < nb; I += nsplit) {
      const double* gI = a.Linv + (long long)I * kBlk * a.ld;
      const int nst = (I + 1) * kStagesPerBlk;
      for (int kt = 0; kt < nst; ++kt) {
        long long tp0 = 0;
        if (a.dbg) tp0 = clock64();
        mbar_wait(empty + slot, phase ^ 1u);
        if (a.dbg && lane == 0 && pw == 0) {
          atomicAdd(a.dbg + 3, (unsigned long long)(clock64() - tp0));  // cycles the producer waited on empty
          atomicAdd(a.dbg + 4, 1ull);
          tp0 = clock64();
        }
        {
          // this warp's quarter of the A tile: rows [pw*32, pw*32+32) x BK/2 chunks of 16 B;
          // lane -> (row = it * RPI + lane / CPRA, chunk = lane % CPRA)
          constexpr int CPRA = BK / 2, RPI = 32 / CPRA, RW = kBlk / kPProdWarps;
          double* s = stA(slot) + (pw * RW) * AS;
          const double* g = gI + kt * BK + (long long)(pw * RW + (lane / CPRA)) * a.ld + (lane % CPRA) * 2;
          double* sd = s + (lane / CPRA) * AS + (lane % CPRA) * 2;
#pragma unroll 8
          for (int it = 0; it < RW / RPI; ++it)
            cp_async16(sd + it * RPI * AS, g + (long long)it * RPI * a.ld, true);
        }
        {
          // this warp's quarter of the K* tile: k rows [pw*8, pw*8+8)
          constexpr int KR = BK / kPProdWarps;
          constexpr int CPR = TQ / 2;             // 16-byte chunks per k row
          double* s = stK(slot) + pw * KR * KS;
          const double* g = gK0 + (long long)(kt * BK + pw * KR) * a.ldks;
#pragma unroll 4
          for (int it = 0; it < KR * CPR / 32; ++it) {
            const int ch = it * 32 + lane, k = ch / CPR, c2 = (ch % CPR) * 2;
            cp_async16(s + k * KS + c2, g + (long long)k * a.ldks + c2, true);
          }
        }
        if (Cfg::GMUL && pw == kPProdWarps - 1) {
          double* s = stG(slot);
          const double* g = a.G + (long long)kt * BK * P;
          for (int ch = lane; ch < BK * P / 2; ch += 32) cp_async16(s + ch * 2, g + ch * 2, true);
        }
        mbar_arrive_cp_async(full + slot);
        if (a.dbg && lane == 0 && pw == 0) atomicAdd(a.dbg + 5, (unsigned long long)(clock64() - tp0));  // cycles to issue a stage
        if (++slot == S) { slot = 0; phase ^= 1u; }
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
  } else {
    // ================= consumers: DMMA on the staged tiles ====================================================
    const int wm = warp >> 2, wn = warp & 3;
    const int lr = lane >> 2, lk = lane & 3;
    double acc[8][NF][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int f = 0; f < NF; ++f) acc[i][f][0] = acc[i][f][1] = 0.0;
    int slot = 0;
    unsigned phase = 0;
    for (int I = split; I < nb; I += nsplit) {
      const int nst = (I + 1) * kStagesPerBlk;
      for (int kt = 0; kt < nst; ++kt) {
        if (a.dbg) {
          const long long t0 = clock64();
          mbar_wait(full + slot, phase);
          const long long dt = clock64() - t0;
          if (lane == 0) {
            atomicAdd(a.dbg + 0, (unsigned long long)dt);             // cycles consumer warps spent waiting on full
            atomicAdd(a.dbg + 1, 1ull);                               // consumer stage count
            if (dt > 300) atomicAdd(a.dbg + 2, 1ull);                 // stages that actually blocked
          }
        } else {
          mbar_wait(full + slot, phase);
        }
        const double* As = stA(slot) + (wm * 64 + lr) * AS + lk;
        const double* Ks = stK(slot) + lk * KS + wn * QW + lr;
        const double* Gs = stG(slot) + lk * P;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; ++k4) {
          double af[8], kq[H], bf[NF];
#pragma unroll
          for (int i = 0; i < 8; ++i) af[i] = As[i * 8 * AS + k4 * 4];
#pragma unroll
          for (int h = 0; h < H; ++h) kq[h] = Ks[k4 * 4 * KS + h * 8];
          if (Cfg::GMUL) {
#pragma unroll
            for (int q = 0; q < P; ++q) {
              const double g = Gs[k4 * 4 * P + q];
#pragma unroll
              for (int h = 0; h < H; ++h) bf[q * H + h] = kq[h] * g;
            }
          } else {
#pragma unroll
            for (int h = 0; h < H; ++h) bf[h] = kq[h];
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int f = 0; f < NF; ++f) dmma884(acc[i][f][0], acc[i][f][1], af[i], bf[f]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
        if (++slot == S) { slot = 0; phase ^= 1u; }
      }
      // per-query Gram of this row block: sum over the warp's 64 rows of V[:,q] V[:,r]
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double sp[NPair];
#pragma unroll
          for (int e = 0; e < NPair; ++e) sp[e] = 0.0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            int e = 0;
#pragma unroll
            for (int q = 0; q < P; ++q)
#pragma unroll
              for (int r = q; r < P; ++r) {
                sp[e] = fma(acc[i][q * H + h][c], acc[i][r * H + h][c], sp[e]);
                ++e;
              }
          }
#pragma unroll
          for (int e = 0; e < NPair; ++e) {
            double v = sp[e];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lr == 0) {
              int t = wn * QW + h * 8 + lk * 2 + c;
              Ssm[(wm * TQ + t) * NPair + e] += v;  // single owner per (wm, t, e): no atomics, fixed order
            }
          }
        }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int f = 0; f < NF; ++f) acc[i][f][0] = acc[i][f][1] = 0.0;
    }
  }
  __syncthreads();
  if (a.dbg && tid == 0) {
    atomicAdd(a.dbg + 6, (unsigned long long)(clock64() - tk0));  // CTA lifetime cycles
    atomicAdd(a.dbg + 7, 1ull);
  }
  for (int i = tid; i < TQ * NPair; i += kPThreads) {
    int t = i / NPair, e = i % NPair;
    double v = Ssm[(0 * TQ + t) * NPair + e] + Ssm[(1 * TQ + t) * NPair + e];
    a.Spart[((long long)split * a.Qpad + q0 + t) * NPair + e] = v;
  }
}

// kb*[i, q] = K*[i, q] * (G[i,:] . UHq[q,:])   (fold-in form, control_affine_model.py:536)
__global__ void fold_kbstar_kernel(const double* __restrict__ Kstar, int ldks, const double* __restrict__ G,
                                   const double* __restrict__ UHq, int p, int Npad, int Q, int Qpad,
                                   double* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int i0 = blockIdx.y * 64;
  if (q >= Qpad) return;
  double uh[BCBF_MAX_P_DIM];
  for (int j = 0; j < p; ++j) uh[j] = (q < Q) ? UHq[(long long)q * p + j] : 0.0;
  for (int i = i0; i < min(Npad, i0 + 64); ++i) {
    double w = 0.0;
    for (int j = 0; j < p; ++j) w = fma(G[(long long)i * p + j], uh[j], w);
    out[(long long)i * Qpad + q] = (q < Q) ? Kstar[(long long)i * ldks + q] * w : 0.0;
  }
}

// Bk[q] = kss * B - sum_split S   (matrix form)   /   svar[q] = kss * uh^T B uh - sum_split S  (fold-in)
__global__ void post_var_finalize_kernel(const double* __restrict__ Spart, int Qpad, int nsplit, int Q, int p,
                                         const double* __restrict__ Bmat, double kss, double* __restrict__ Bk,
                                         const double* __restrict__ UHq, double* __restrict__ svar) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  if (Bk) {
    const int npair = p * (p + 1) / 2;
    int e = 0;
    for (int i = 0; i < p; ++i)
      for (int j = i; j < p; ++j) {
        double s = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) s += Spart[((long long)sp * Qpad + q) * npair + e];
        double v = kss * Bmat[i * p + j] - s;
        Bk[((long long)q * p + i) * p + j] = v;
        Bk[((long long)q * p + j) * p + i] = v;
        ++e;
      }
  } else {
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += Spart[(long long)sp * Qpad + q];
    double ubu = 0.0;
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < p; ++j) ubu += UHq[(long long)q * p + i] * Bmat[i * p + j] * UHq[(long long)q * p + j];
    svar[q] = kss * ubu - s;
  }
}

// Partial posterior mean: part[split][q][c] = sum_{i in split} Kstar[i][q] * W[i][c]
constexpr int kMeanThreads = 128;
constexpr int kMeanRows = 64;
__global__ void __launch_bounds__(kMeanThreads)
post_mean_partial_kernel(const double* __restrict__ Kstar, int ldks, const double* __restrict__ W, int nc, int N,
                         int Q, int rows_per_split, double* __restrict__ part, int Qpad) {
  __shared__ double Wsm[kMeanRows][BCBF_MAX_N_DIM * BCBF_MAX_P_DIM];
  const int q = blockIdx.x * kMeanThreads + threadIdx.x;
  const int i0 = blockIdx.y * rows_per_split, i1 = min(N, i0 + rows_per_split);
  double acc[BCBF_MAX_N_DIM * BCBF_MAX_P_DIM];
#pragma unroll
  for (int c = 0; c < BCBF_MAX_N_DIM * BCBF_MAX_P_DIM; ++c) acc[c] = 0.0;
  for (int ib = i0; ib < i1; ib += kMeanRows) {
    const int rows = min(kMeanRows, i1 - ib);
    __syncthreads();
    for (int idx = threadIdx.x; idx < rows * nc; idx += kMeanThreads) Wsm[idx / nc][idx % nc] = W[(long long)(ib + idx / nc) * nc + idx % nc];
    __syncthreads();
    if (q < Q) {
      // 8 independent K* loads in flight per thread (the loop is latency-bound otherwise)
      int r = 0;
      for (; r + 8 <= rows; r += 8) {
        double kv8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) kv8[u] = Kstar[(long long)(ib + r + u) * ldks + q];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int c = 0; c < BCBF_MAX_N_DIM * BCBF_MAX_P_DIM; ++c)
            if (c < nc) acc[c] = fma(kv8[u], Wsm[r + u][c], acc[c]);
      }
      for (; r < rows; ++r) {
        const double kvv = Kstar[(long long)(ib + r) * ldks + q];
#pragma unroll
        for (int c = 0; c < BCBF_MAX_N_DIM * BCBF_MAX_P_DIM; ++c)
          if (c < nc) acc[c] = fma(kvv, Wsm[r][c], acc[c]);
      }
    }
  }
  if (q < Q)
    for (int c = 0; c < nc; ++c) part[((long long)blockIdx.y * Qpad + q) * nc + c] = acc[c];
}

__global__ void post_mean_finalize_kernel(const double* __restrict__ part, int Qpad, int nsplit, int Q, int nc,
                                          const double* __restrict__ Ct, double* __restrict__ Mk) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Q * nc) return;
  int q = (int)(idx / nc), c = (int)(idx % nc);
  double s = Ct[c];
  for (int sp = 0; sp < nsplit; ++sp) s += part[((long long)sp * Qpad + q) * nc + c];
  Mk[idx] = s;
}

__global__ void contract_u_kernel(const double* __restrict__ Mk, const double* __restrict__ Bk,
                                  const double* __restrict__ UHq, int n, int p, int Q, double* __restrict__ mean,
                                  double* __restrict__ svar) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  double uh[BCBF_MAX_P_DIM];
  for (int j = 0; j < p; ++j) uh[j] = UHq[(long long)q * p + j];
  if (mean && Mk)
    for (int r = 0; r < n; ++r) {
      double s = 0.0;
      for (int j = 0; j < p; ++j) s = fma(Mk[((long long)q * n + r) * p + j], uh[j], s);
      mean[(long long)q * n + r] = s;
    }
  if (svar && Bk) {
    double s = 0.0;
    for (int i = 0; i < p; ++i)
      for (int j = 0; j < p; ++j) s = fma(uh[i] * Bk[((long long)q * p + i) * p + j], uh[j], s);
    svar[q] = s;
  }
}

// ---- workspace (per device, grown on demand, freed at process exit) ---------------------------------
struct Workspace {
  double* ptr = nullptr;
  size_t bytes = 0;
};
static Workspace g_ws[3][64];  // slot 0: Gram partials, 1: mean partials, 2: fold-in kb*

static int get_workspace(int slot, size_t bytes, double** out) {
  int dev = 0;
  BCBF_CUDA(cudaGetDevice(&dev));
  Workspace& w = g_ws[slot][dev & 63];
  if (w.bytes < bytes) {
    if (w.ptr) BCBF_CUDA(cudaFree(w.ptr));
    w.ptr = nullptr;
    w.bytes = 0;
    BCBF_CUDA(cudaMalloc(&w.ptr, bytes));
    w.bytes = bytes;
  }
  *out = w.ptr;
  return BCBF_OK;
}

// ---- optional event profiling of the dominant kernel (bench.py's roofline leg) ------------------------
struct KernelProfile {
  bool on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
};
static KernelProfile g_prof;

static unsigned long long* g_dbg = nullptr;  // 8 counters, enabled by bcbf_debug_counters(1, ...)

template <class Cfg>
static int launch_post_var(PostArgs a, cudaStream_t stream) {
  a.dbg = g_dbg;
  BCBF_CUDA(cudaFuncSetAttribute(post_var_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SmemBytes));
  dim3 grid(a.Qpad / Cfg::TQ, a.nsplit);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_prof.on) {
    BCBF_CUDA(cudaEventCreate(&e0));
    BCBF_CUDA(cudaEventCreate(&e1));
    BCBF_CUDA(cudaEventRecord(e0, stream));
  }
  post_var_kernel<Cfg><<<grid, kPThreads, Cfg::SmemBytes, stream>>>(a);
  BCBF_LAUNCH_CHECK();
  if (g_prof.on) {
    BCBF_CUDA(cudaEventRecord(e1, stream));
    g_prof.spans.emplace_back(e0, e1);
  }
  return BCBF_OK;
}

static int pick_split(int tiles, int nb) {
  // fill the 148 SMs when there are few query tiles; keep whole-Linv walks per CTA otherwise
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (tiles >= sms) return 1;
  int s = (sms + tiles - 1) / tiles;
  if (s > nb) s = nb;
  return s < 1 ? 1 : s;
}

static int var_dispatch(bool fold, int p, PostArgs& a, const double* UHq, cudaStream_t stream) {
  // queries per CTA: p=1 / fold-in: 96, p=2: 64, p=3,4: 32.  Fills a.Qpad / a.nsplit / a.Spart.
  const int TQ = (fold || p == 1) ? 96 : (p == 2 ? 64 : 32);
  a.Qpad = ((a.Q + TQ - 1) / TQ) * TQ;
  BCBF_REQUIRE(a.ldks >= a.Qpad,
               "posterior: Kstar leading dimension %d < padded query count %d (pad to a multiple of 192)", a.ldks,
               a.Qpad);
  a.nsplit = pick_split(a.Qpad / TQ, a.Npad / kBlk);
  const int npair = fold ? 1 : p * (p + 1) / 2;
  double* ws = nullptr;
  int rc = get_workspace(0, sizeof(double) * (size_t)a.nsplit * a.Qpad * npair, &ws);
  if (rc != BCBF_OK) return rc;
  a.Spart = ws;
  if (fold) {
    // kb* = K* (.) (G UHq^T) once, then the p = 1 contraction on it
    double* kb = nullptr;
    rc = get_workspace(2, sizeof(double) * (size_t)a.Npad * a.Qpad, &kb);
    if (rc != BCBF_OK) return rc;
    fold_kbstar_kernel<<<dim3(ceil_div(a.Qpad, 128), ceil_div(a.Npad, 64)), 128, 0, stream>>>(
        a.Kstar, a.ldks, a.G, UHq, p, a.Npad, a.Q, a.Qpad, kb);
    BCBF_LAUNCH_CHECK();
    a.Kstar = kb;
    a.ldks = a.Qpad;
    return launch_post_var<PostCfg<1, 3, false>>(a, stream);
  }
  switch (p) {
    case 1: return launch_post_var<PostCfg<1, 3, true>>(a, stream);
    case 2: return launch_post_var<PostCfg<2, 2, true>>(a, stream);
    case 3: return launch_post_var<PostCfg<3, 1, true>>(a, stream);
    case 4: return launch_post_var<PostCfg<4, 1, true>>(a, stream);
    default: break;
  }
  set_last_error("posterior: p=%d unsupported", p);
  return BCBF_ERR_INVALID;
}

}  // namespace bcbf

using namespace bcbf;

static int run_mean(const double* Kstar, int ldks, const double* W, const double* Ct, int N, int n, int p, int Q,
                    double* Mk, cudaStream_t stream) {
  const int nc = n * p;
  const int qblocks = ceil_div(Q, kMeanThreads);
  int nsplit = (148 * 16) / qblocks;  // ~16 CTAs of 128 threads per SM: enough loads in flight to stream K* at HBM rate
  if (nsplit < 1) nsplit = 1;
  int max_split = ceil_div(N, kMeanRows);
  if (nsplit > max_split) nsplit = max_split;
  const int rows_per_split = ceil_div(ceil_div(N, nsplit), kMeanRows) * kMeanRows;
  nsplit = ceil_div(N, rows_per_split);
  const int Qpad = qblocks * kMeanThreads;
  double* ws = nullptr;
  {
    int rc = get_workspace(1, sizeof(double) * (size_t)nsplit * Qpad * nc, &ws);
    if (rc != BCBF_OK) return rc;
  }
  post_mean_partial_kernel<<<dim3(qblocks, nsplit), kMeanThreads, 0, stream>>>(Kstar, ldks, W, nc, N, Q,
                                                                               rows_per_split, ws, Qpad);
  BCBF_LAUNCH_CHECK();
  post_mean_finalize_kernel<<<ceil_div((long long)Q * nc, 256), 256, 0, stream>>>(ws, Qpad, nsplit, Q, nc, Ct, Mk);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_posterior_blocks(const double* Linv, int ld, int Npad, const double* Kstar, int ldks,
                                     const double* G, const double* W, const double* Bmat, const double* Ct,
                                     double kss, int n, int p, int Q, double* Mk, double* Bk, void* stream_) {
  ::bcbf::ScratchScope scratch_scope(static_cast<cudaStream_t>(stream_));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(Linv && Kstar && G && Bmat, "bcbf_posterior_blocks: null pointer");
  BCBF_REQUIRE(n >= 1 && n <= BCBF_MAX_N_DIM && p >= 1 && p <= BCBF_MAX_P_DIM, "bcbf_posterior_blocks: n=%d p=%d", n, p);
  BCBF_REQUIRE(Npad > 0 && Npad % kBlk == 0 && ld >= Npad && ld % 2 == 0 && Q >= 1 && ldks % 2 == 0,
               "bcbf_posterior_blocks: Npad=%d ld=%d Q=%d ldks=%d", Npad, ld, Q, ldks);
  if (Mk) {
    BCBF_REQUIRE(W && Ct, "bcbf_posterior_blocks: Mk requested without W / Ct");
    int rc = run_mean(Kstar, ldks, W, Ct, Npad, n, p, Q, Mk, stream);
    if (rc != BCBF_OK) return rc;
  }
  if (Bk) {
    PostArgs a{};
    a.Linv = Linv; a.ld = ld; a.Npad = Npad; a.Kstar = Kstar; a.ldks = ldks; a.G = G; a.Q = Q;
    int rc = var_dispatch(false, p, a, nullptr, stream);
    if (rc != BCBF_OK) return rc;
    post_var_finalize_kernel<<<ceil_div(Q, 128), 128, 0, stream>>>(a.Spart, a.Qpad, a.nsplit, Q, p, Bmat, kss, Bk,
                                                                   nullptr, nullptr);
    BCBF_LAUNCH_CHECK();
  }
  return BCBF_OK;
}

extern "C" int bcbf_posterior_fu(const double* Linv, int ld, int Npad, const double* Kstar, int ldks, const double* G,
                                 const double* alpha, const double* Bmat, const double* C, const double* UHq, double kss,
                                 int n, int p, int Q, double* mean, double* svar, void* stream_) {
  ::bcbf::ScratchScope scratch_scope(static_cast<cudaStream_t>(stream_));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  (void)alpha; (void)C; (void)mean;
  BCBF_REQUIRE(Linv && Kstar && G && Bmat && UHq && svar, "bcbf_posterior_fu: null pointer");
  BCBF_REQUIRE(n >= 1 && n <= BCBF_MAX_N_DIM && p >= 1 && p <= BCBF_MAX_P_DIM, "bcbf_posterior_fu: n=%d p=%d", n, p);
  BCBF_REQUIRE(Npad > 0 && Npad % kBlk == 0 && ld >= Npad && ld % 2 == 0 && Q >= 1 && ldks % 2 == 0,
               "bcbf_posterior_fu: Npad=%d ld=%d Q=%d ldks=%d", Npad, ld, Q, ldks);
  PostArgs a{};
  a.Linv = Linv; a.ld = ld; a.Npad = Npad; a.Kstar = Kstar; a.ldks = ldks; a.G = G; a.Q = Q;
  int rc = var_dispatch(true, p, a, UHq, stream);
  if (rc != BCBF_OK) return rc;
  post_var_finalize_kernel<<<ceil_div(Q, 128), 128, 0, stream>>>(a.Spart, a.Qpad, a.nsplit, Q, p, Bmat, kss, nullptr,
                                                                 UHq, svar);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_contract_u(const double* Mk, const double* Bk, const double* UHq, int n, int p, int Q,
                               double* mean, double* svar, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(UHq && Q >= 1 && n >= 1 && p >= 1 && p <= BCBF_MAX_P_DIM, "bcbf_contract_u: bad arguments");
  contract_u_kernel<<<ceil_div(Q, 128), 128, 0, stream>>>(Mk, Bk, UHq, n, p, Q, mean, svar);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

// Event profiling of post_var_kernel launches: enable (clears history), then read after the timed region.
extern "C" int bcbf_profile_enable(int on) {
  for (auto& sp : g_prof.spans) {
    cudaEventDestroy(sp.first);
    cudaEventDestroy(sp.second);
  }
  g_prof.spans.clear();
  g_prof.on = on != 0;
  return BCBF_OK;
}

extern "C" int bcbf_profile_read(double* total_ms, int* launches) {
  double tot = 0.0;
  for (auto& sp : g_prof.spans) {
    BCBF_CUDA(cudaEventSynchronize(sp.second));
    float ms = 0.f;
    BCBF_CUDA(cudaEventElapsedTime(&ms, sp.first, sp.second));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = (int)g_prof.spans.size();
  return BCBF_OK;
}

// Pipeline counters of post_var_kernel (development aid; adds clock64/atomics to the kernel while enabled):
//   [0] cycles consumer warps waited on `full`  [1] consumer (warp, stage) count  [2] waits > 300 cycles
//   [3] cycles the producer waited on `empty`   [4] producer stage count          [5] cycles spent issuing copies
//   [6] CTA lifetime cycles (sum)               [7] CTA count
extern "C" int bcbf_debug_counters(int enable, unsigned long long out[8]) {
  if (out != nullptr && g_dbg != nullptr) {
    BCBF_CUDA(cudaDeviceSynchronize());
    BCBF_CUDA(cudaMemcpy(out, g_dbg, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  if (enable && g_dbg == nullptr) {
    BCBF_CUDA(cudaMalloc(&g_dbg, 8 * sizeof(unsigned long long)));
  }
  if (g_dbg != nullptr) BCBF_CUDA(cudaMemset(g_dbg, 0, 8 * sizeof(unsigned long long)));
  if (!enable && g_dbg != nullptr) {
    BCBF_CUDA(cudaFree(g_dbg));
    g_dbg = nullptr;
  }
  return BCBF_OK;
}
