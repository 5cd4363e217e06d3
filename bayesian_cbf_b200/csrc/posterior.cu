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

constexpr int kPA_Stride = 20;                  // A tile row stride (16 + 4 pad doubles)
constexpr int kPA_Elems = 128 * kPA_Stride;     // 2560
constexpr int kPStages = 3;
constexpr int kPThreads = 256;

template <int P_, int QW_, bool FOLD_>
struct PostCfg {
  static constexpr int P = P_;                       // columns per query (p, or 1 when u is folded in)
  static constexpr int QW = QW_;                     // queries per warp column (multiple of 8)
  static constexpr bool FOLD = FOLD_;
  static constexpr int TQ = 2 * QW;                  // queries per CTA
  static constexpr int BN = P * TQ;                  // tile columns
  static constexpr int H = QW / 8;                   // n-fragments per (warp, q)
  static constexpr int NF = P * H;                   // n-fragments per warp
  static constexpr int BStride = BN + 4;             // == 4 (mod 16): conflict-free fragment reads
  static constexpr int BElems = 16 * BStride;
  static constexpr int NPair = P * (P + 1) / 2;
  static constexpr int StageElems = kPA_Elems + BElems;
  static constexpr int NChunk = (8 * TQ + kPThreads - 1) / kPThreads;  // double2 Kstar chunks per thread/stage
  static constexpr int SmemBytes = (kPStages * StageElems + 4 * TQ * NPair) * (int)sizeof(double);
  static_assert(BN % 16 == 0, "tile width must keep the padded stride at 4 mod 16");
  static_assert(QW % 8 == 0, "");
};

struct PostArgs {
  const double* Linv; int ld; int Npad;
  const double* Kstar; int ldks;
  const double* G;       // (Npad, pg) row-major, pad rows zero
  int pg;                // true p (row length of G / UHq)
  const double* UHq;     // (Q, pg) fold-in mode only
  int Q;
  double* Spart;         // [nsplit][Qpad][NPair] partial Gram sums
  int Qpad; int nsplit;
};

template <class Cfg>
__global__ void __launch_bounds__(kPThreads, 1) post_var_kernel(PostArgs a) {
  constexpr int P = Cfg::P, QW = Cfg::QW, TQ = Cfg::TQ, H = Cfg::H, NF = Cfg::NF, BS = Cfg::BStride;
  constexpr int NPair = Cfg::NPair, NCH = Cfg::NChunk;
  extern __shared__ __align__(16) double smem[];
  double* Ssm = smem + kPStages * Cfg::StageElems;  // [4][TQ][NPair]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int lr = lane >> 2, lk = lane & 3;
  const int q0 = blockIdx.x * TQ;
  const int nb = a.Npad / kBlk;
  const int split = blockIdx.y, nsplit = a.nsplit;

  for (int i = tid; i < 4 * TQ * NPair; i += kPThreads) Ssm[i] = 0.0;

  // ---- iteration space: row blocks I = split, split+nsplit, ... ; stages kt in [0, (I+1)*8) ----------
  auto block_stages = [](int I) { return (I + 1) * (kBlk / 16); };
  long long total = 0;
  for (int I = split; I < nb; I += nsplit) total += block_stages(I);

  struct Cursor { int I, kt; };
  auto advance = [&](Cursor& c) {
    if (++c.kt == block_stages(c.I)) { c.kt = 0; c.I += nsplit; }
  };
  auto stA = [&](int s) { return smem + s * Cfg::StageElems; };
  auto stB = [&](int s) { return smem + s * Cfg::StageElems + kPA_Elems; };

  auto loadA = [&](const Cursor& c, int slot) {
    const double* g = a.Linv + (long long)c.I * kBlk * a.ld + c.kt * 16;
    double* s = stA(slot);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int ch = tid + i * kPThreads;
      int row = ch >> 3, kc = (ch & 7) * 2;
      cp_async16(s + row * kPA_Stride + kc, g + (long long)row * a.ld + kc, true);
    }
  };

  // B-operand producer: registers <- Kstar / G (global), then registers -> shared tile
  double2 kv[NCH];
  double gq[NCH][Cfg::FOLD ? 1 : P];
  auto ldgB = [&](const Cursor& c) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      int ch = tid + i * kPThreads;
      if (ch < 8 * TQ) {
        int k = ch / (TQ / 2), t2 = (ch % (TQ / 2)) * 2;
        int row = c.kt * 16 + k;
        kv[i] = *reinterpret_cast<const double2*>(a.Kstar + (long long)row * a.ldks + q0 + t2);
        if (Cfg::FOLD) {
          // (G[row] . uh[t]) for the two queries of this chunk is folded into kv directly
          double g0 = 0.0, g1 = 0.0;
          for (int j = 0; j < a.pg; ++j) {
            double g = a.G[(long long)row * a.pg + j];
            int qa = min(q0 + t2, a.Q - 1), qb = min(q0 + t2 + 1, a.Q - 1);
            g0 = fma(g, a.UHq[(long long)qa * a.pg + j], g0);
            g1 = fma(g, a.UHq[(long long)qb * a.pg + j], g1);
          }
          kv[i].x *= g0;
          kv[i].y *= g1;
          gq[i][0] = 1.0;
        } else {
#pragma unroll
          for (int q = 0; q < P; ++q) gq[i][q] = a.G[(long long)row * P + q];
        }
      }
    }
  };
  auto stsB = [&](int slot) {
    double* s = stB(slot);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      int ch = tid + i * kPThreads;
      if (ch < 8 * TQ) {
        int k = ch / (TQ / 2), t2 = (ch % (TQ / 2)) * 2;
#pragma unroll
        for (int q = 0; q < P; ++q) {
          double g = Cfg::FOLD ? 1.0 : gq[i][q];
          *reinterpret_cast<double2*>(s + k * BS + q * TQ + t2) = make_double2(kv[i].x * g, kv[i].y * g);
        }
      }
    }
  };

  double acc[4][NF][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int f = 0; f < NF; ++f) acc[i][f][0] = acc[i][f][1] = 0.0;

  Cursor cl{split, 0};  // loader cursor (A, runs 2 stages ahead)
  Cursor cb{split, 0};  // B producer cursor
  Cursor cc{split, 0};  // consumer cursor
  // prologue
  if (total > 0) { loadA(cl, 0); advance(cl); }
  cp_async_commit();
  if (total > 1) { loadA(cl, 1); advance(cl); }
  cp_async_commit();
  if (total > 0) { ldgB(cb); advance(cb); stsB(0); }
  if (total > 1) { ldgB(cb); advance(cb); }

  for (long long s = 0; s < total; ++s) {
    cp_async_wait<1>();
    __syncthreads();
    const int slot = (int)(s % kPStages);
    if (s + 2 < total) { loadA(cl, (int)((s + 2) % kPStages)); advance(cl); }
    cp_async_commit();
    if (s + 1 < total) stsB((int)((s + 1) % kPStages));
    if (s + 2 < total) { ldgB(cb); advance(cb); }

    const double* As = stA(slot);
    const double* Bs = stB(slot);
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      const int kk = k4 * 4 + lk;
      double af[4], bf[NF];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = As[(wm * 32 + i * 8 + lr) * kPA_Stride + kk];
#pragma unroll
      for (int q = 0; q < P; ++q)
#pragma unroll
        for (int h = 0; h < H; ++h) bf[q * H + h] = Bs[kk * BS + q * TQ + wn * QW + h * 8 + lr];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int f = 0; f < NF; ++f) dmma884(acc[i][f][0], acc[i][f][1], af[i], bf[f]);
    }

    const bool last_of_block = (cc.kt + 1 == block_stages(cc.I));
    if (last_of_block) {
      // per-query Gram of this row block: sum over the warp's 32 rows of V[:,q] V[:,r]
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          double sp[NPair];
#pragma unroll
          for (int e = 0; e < NPair; ++e) sp[e] = 0.0;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
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
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int f = 0; f < NF; ++f) acc[i][f][0] = acc[i][f][1] = 0.0;
    }
    advance(cc);
  }
  cp_async_wait<0>();
  __syncthreads();
  for (int i = tid; i < TQ * NPair; i += kPThreads) {
    int t = i / NPair, e = i % NPair;
    double v = Ssm[(0 * TQ + t) * NPair + e] + Ssm[(1 * TQ + t) * NPair + e] + Ssm[(2 * TQ + t) * NPair + e] +
               Ssm[(3 * TQ + t) * NPair + e];
    a.Spart[((long long)split * a.Qpad + q0 + t) * NPair + e] = v;
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
      for (int r = 0; r < rows; ++r) {
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
static Workspace g_ws[64];

static int get_workspace(size_t bytes, double** out) {
  int dev = 0;
  BCBF_CUDA(cudaGetDevice(&dev));
  Workspace& w = g_ws[dev & 63];
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

template <class Cfg>
static int launch_post_var(PostArgs a, cudaStream_t stream) {
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

static int var_dispatch(bool fold, int p, PostArgs& a, cudaStream_t stream) {
  // (P, QW) -> TQ: p=1:96  p=2:48  p=3:32  p=4:32 ; fold-in: 96.  Fills a.Qpad / a.nsplit / a.Spart.
  const int TQ = (fold || p == 1) ? 96 : (p == 2 ? 48 : 32);
  a.Qpad = ((a.Q + TQ - 1) / TQ) * TQ;
  BCBF_REQUIRE(a.ldks >= a.Qpad,
               "posterior: Kstar leading dimension %d < padded query count %d (pad to a multiple of 96)", a.ldks,
               a.Qpad);
  a.nsplit = pick_split(a.Qpad / TQ, a.Npad / kBlk);
  const int npair = fold ? 1 : p * (p + 1) / 2;
  double* ws = nullptr;
  int rc = get_workspace(sizeof(double) * (size_t)a.nsplit * a.Qpad * npair, &ws);
  if (rc != BCBF_OK) return rc;
  a.Spart = ws;
  if (fold) return launch_post_var<PostCfg<1, 48, true>>(a, stream);
  switch (p) {
    case 1: return launch_post_var<PostCfg<1, 48, false>>(a, stream);
    case 2: return launch_post_var<PostCfg<2, 24, false>>(a, stream);
    case 3: return launch_post_var<PostCfg<3, 16, false>>(a, stream);
    case 4: return launch_post_var<PostCfg<4, 16, false>>(a, stream);
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
  int nsplit = 296 / qblocks;
  if (nsplit < 1) nsplit = 1;
  int max_split = ceil_div(N, kMeanRows);
  if (nsplit > max_split) nsplit = max_split;
  const int rows_per_split = ceil_div(ceil_div(N, nsplit), kMeanRows) * kMeanRows;
  nsplit = ceil_div(N, rows_per_split);
  const int Qpad = qblocks * kMeanThreads;
  double* ws = nullptr;
  // the variance path owns the front of the workspace; the mean partials live in their own allocation
  static double* mean_ws[64] = {nullptr};
  static size_t mean_ws_bytes[64] = {0};
  int dev = 0;
  BCBF_CUDA(cudaGetDevice(&dev));
  size_t need = sizeof(double) * (size_t)nsplit * Qpad * nc;
  if (mean_ws_bytes[dev & 63] < need) {
    if (mean_ws[dev & 63]) BCBF_CUDA(cudaFree(mean_ws[dev & 63]));
    mean_ws[dev & 63] = nullptr;
    mean_ws_bytes[dev & 63] = 0;
    BCBF_CUDA(cudaMalloc(&mean_ws[dev & 63], need));
    mean_ws_bytes[dev & 63] = need;
  }
  ws = mean_ws[dev & 63];
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
    a.Linv = Linv; a.ld = ld; a.Npad = Npad; a.Kstar = Kstar; a.ldks = ldks; a.G = G; a.pg = p; a.UHq = nullptr; a.Q = Q;
    int rc = var_dispatch(false, p, a, stream);
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
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  (void)alpha; (void)C; (void)mean;
  BCBF_REQUIRE(Linv && Kstar && G && Bmat && UHq && svar, "bcbf_posterior_fu: null pointer");
  BCBF_REQUIRE(n >= 1 && n <= BCBF_MAX_N_DIM && p >= 1 && p <= BCBF_MAX_P_DIM, "bcbf_posterior_fu: n=%d p=%d", n, p);
  BCBF_REQUIRE(Npad > 0 && Npad % kBlk == 0 && ld >= Npad && ld % 2 == 0 && Q >= 1 && ldks % 2 == 0,
               "bcbf_posterior_fu: Npad=%d ld=%d Q=%d ldks=%d", Npad, ld, Q, ldks);
  PostArgs a{};
  a.Linv = Linv; a.ld = ld; a.Npad = Npad; a.Kstar = Kstar; a.ldks = ldks; a.G = G; a.pg = p; a.UHq = UHq; a.Q = Q;
  int rc = var_dispatch(true, p, a, stream);
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
