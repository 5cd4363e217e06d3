// FP64-accurate posterior covariance contraction on the int8 tensor cores (tcgen05, sm_100a).
//
//   S(x) = frakB(x)^T Kb^-1 frakB(x) = V^T V,   V = L^-1 frakB(x),   frakB[i,(q,t)] = K*[i,q] G[i,t]
//   (control_affine_model.py:1051-1088 of the reference; the N^2 p flops per query of SURVEY 8d)
//
// The FP64 tensor pipe (DMMA) tops out at 37 TFLOP/s, the int8 pipe at 4.5 POP/s.  Both FP64 operands are therefore split
// error-free into S = 7 signed 8-bit digits in base 256 (an Ozaki-type splitting):
//     L^-1[i,k]  = 2^ea_i * sum_j a_j[i,k] 256^-(j+1),      frakB[k,c] = 2^eb_c * sum_j b_j[k,c] 256^-(j+1),
// with one power-of-two scale per ROW of L^-1 and per COLUMN of frakB, so every digit product a_j b_l is an exact integer
// and every int32 accumulation over k is exact (|sum| <= 7 * 16384 * 2^14 < 2^31).  Products with j + l = d carry the same
// weight 256^-(d+2) and share one accumulator; the 28 products with d <= 6 are kept (the dropped ones are below 2^-56 of
// row-scale x column-scale, the same order as the FP64 rounding of a plain DGEMM; tests/test_gpu_ozaki.py).  The FP64
// value is recombined from the seven int32 accumulators in the epilogue, where the per-query p x p Gram V^T V is formed:
// V never exists in HBM.
//
// Kernel structure (oz_var_kernel): persistent, one CTA per SM, tile = 128 rows of L^-1 x 64 frakB columns (21 queries
// at p = 3), seven 64-column int32 accumulators = 448 of the 512 TMEM columns.  Warp roles: warp 4 streams the operand
// digits with bulk asynchronous copies (cp.async.bulk, one 28 KB + one 14 KB copy per 32-deep K step, 5 stages, the
// digit arrays are stored in HBM exactly in the shared-memory image the MMA wants: un-swizzled K-major core matrices);
// one thread of warp 5 issues the MMAs: digit slice a of L^-1 against slices 0..6-a of frakB CONCATENATED along N (the
// B slices are adjacent in shared memory and the accumulators of consecutive diagonals are adjacent in TMEM), 10 MMAs per
// K step instead of 28, which keeps the shared-memory operand reads under the 128 B/clk limit (measured: 920 cycles per
// K step against a tensor-pipe floor of 896, tools/microbench/umma_i8_probe.cu); warps 0-3 drain TMEM, recombine in
// FP64, apply the scales and reduce the Gram over the 128 rows.
//
// The same digit arithmetic, MMA pattern and pipeline serve two more kernels further down: oz_gemm_kernel (general
// C = alpha A B with an optional triangular operand and inner-dimension balancing; the large levels of bcbf_trtri) and
// oz_update_kernel (rank-K update C += alpha P P^T, one short-lived CTA per tile; the trailing updates of bcbf_potrf).
// oracle/ozaki_oracle.py restates the arithmetic on the CPU with exact integers; all three kernels reproduce it bit for bit
// (tests/test_gpu_ozaki.py).
#include "../../include/bcbf.h"
#include "common.cuh"
#include "tc5.cuh"

namespace bcbf {
namespace oz {

using namespace tc5;

constexpr int S = 7;               // digits per operand
constexpr int TM = 128, TN = 64;   // tile rows (L^-1) x columns (frakB)
constexpr int KSTEP = 32;          // K extent of one int8 MMA
constexpr int A_STEP = S * TM * KSTEP;  // 28672 bytes of L^-1 digits per K step
constexpr int B_STEP = S * TN * KSTEP;  // 14336 bytes of frakB digits per K step
constexpr int STAGE = A_STEP + B_STEP;
constexpr int NSTAGE = 5;
constexpr int kMaxNpad = 18432;    // 7 * Npad * 2^14 < 2^31
constexpr int kEpiThreads = 128;
constexpr int kThreads = 192;
constexpr int kSmemBytes = NSTAGE * STAGE + 128 + TN * 8 + 4 * 160 * 8;

// power-of-two scale 2^e with |x| / 2^e <= 0.498 for all |x| <= mx (balanced base-256 digits span (-0.502, 0.498))
__device__ __forceinline__ double scale_of(double mx) {
  if (!(mx > 0.0)) return 1.0;
  int ex;
  const double f = frexp(mx, &ex);  // mx = f 2^ex, f in [0.5, 1)
  int e = ex + 1;
  if (f * 0.5 >= 0.498) e += 1;
  return ldexp(1.0, e);
}

// x (|x| <= 0.498) -> SD signed digits, x ~= sum_j d[j] 256^-(j+1)  (SD = 7: 56 bits; SD = 6: 48 bits)
template <int SD>
__device__ __forceinline__ void digits_of(double x, int (&d)[SD]) {
  long long I = __double2ll_rn(x * __longlong_as_double((1023LL + 8 * SD) << 52));  // x 2^(8 SD), one rounding
#pragma unroll
  for (int j = SD - 1; j >= 1; --j) {
    const int b = static_cast<int>(static_cast<signed char>(I & 0xFF));
    d[j] = b;
    I = (I - b) >> 8;
  }
  d[0] = static_cast<int>(I < -128 ? -128 : (I > 127 ? 127 : I));
}

// ---- L^-1 -> digit blobs (once per fit) -------------------------------------------------------------------------
__global__ void rowscale_kernel(const double* __restrict__ Linv, int ld, int Npad, double* __restrict__ rowscale) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= Npad) return;
  double mx = 0.0;
  for (int k = lane; k <= row; k += 32) mx = fmax(mx, fabs(Linv[(long long)row * ld + k]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) rowscale[row] = scale_of(mx);
}

// blob of row block I, K step ks (ks < 4 (I+1)):  [slice SD][row group 16][k chunk 2][row 8][16 bytes]
template <int SD>
__global__ void __launch_bounds__(256) split_factor_kernel(const double* __restrict__ Linv, int ld,
                                                           const double* __restrict__ rowscale,
                                                           int8_t* __restrict__ blob) {
  constexpr int S = SD, A_STEP = SD * TM * KSTEP;   // (shadow the 7-digit constants of the GEMM kernels)
  const int kc = blockIdx.x, I = blockIdx.y;
  if (kc > I) return;
  const int r = threadIdx.x % TM, half = threadIdx.x / TM;
  const long long row = (long long)I * TM + r;
  const double inv = 1.0 / rowscale[row];
  int8_t* base = blob + (2LL * I * (I + 1) + 4LL * kc) * A_STEP;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int c16 = half * 4 + c;  // 16-wide k chunk of this 128-wide block column
    const double* src = Linv + row * ld + (long long)kc * TM + c16 * 16;
    uint32_t w[S][4];
#pragma unroll
    for (int s = 0; s < S; ++s) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      int d[S];
      digits_of(src[k] * inv, d);
#pragma unroll
      for (int s = 0; s < S; ++s) w[s][k / 4] |= static_cast<uint32_t>(d[s] & 0xFF) << (8 * (k % 4));
    }
    int8_t* dst = base + (long long)(c16 / 2) * A_STEP + (r / 8) * 256 + (c16 % 2) * 128 + (r % 8) * 16;
#pragma unroll
    for (int s = 0; s < S; ++s)
      *reinterpret_cast<uint4*>(dst + s * (TM * KSTEP)) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
  }
}

// ---- frakB -> digit blobs (per query batch) -----------------------------------------------------------------------
// colmax[q*p + t] = max_i |K*[i,q] G[i,t]|  (as the bit pattern of a non-negative double: unsigned order = value order)
// and, in the same pass over K*, the partial posterior mean  part[split][q][c] = sum_{i in split} K*[i,q] W[i,c]
// (M_k = C^T + K*^T W, control_affine_model.py:1079-1088); W may be NULL (covariance only).
constexpr int kCmRows = 256;
template <int P, int NCMAX>
__global__ void __launch_bounds__(128) colmax_mean_kernel(const double* __restrict__ Kstar, int ldks,
                                                          const double* __restrict__ G, const double* __restrict__ W,
                                                          int nc, int Npad, int Q, int Qpad,
                                                          unsigned long long* __restrict__ colmax,
                                                          double* __restrict__ part) {
  __shared__ double Gs[kCmRows * P];
  __shared__ double Ws[64 * NCMAX];
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int i0 = blockIdx.y * kCmRows;
  const int rows = min(Npad - i0, kCmRows);
  for (int x = threadIdx.x; x < rows * P; x += blockDim.x) Gs[x] = fabs(G[(long long)i0 * P + x]);
  double mx[P];
#pragma unroll
  for (int t = 0; t < P; ++t) mx[t] = 0.0;
  double acc[NCMAX];
#pragma unroll
  for (int c = 0; c < NCMAX; ++c) acc[c] = 0.0;
  for (int ib = 0; ib < rows; ib += 64) {
    const int rb = min(64, rows - ib);
    __syncthreads();
    if (W)
      for (int x = threadIdx.x; x < rb * NCMAX; x += blockDim.x) {
        const int r = x / NCMAX, c = x % NCMAX;
        Ws[x] = c < nc ? W[(long long)(i0 + ib + r) * nc + c] : 0.0;
      }
    __syncthreads();
    if (q < Q) {
      for (int r = 0; r < rb; r += 8) {  // rows is a multiple of 128
        double kv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) kv[u] = Kstar[(long long)(i0 + ib + r + u) * ldks + q];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double ka = fabs(kv[u]);
#pragma unroll
          for (int t = 0; t < P; ++t) mx[t] = fmax(mx[t], ka * Gs[(ib + r + u) * P + t]);
          if (W) {
#pragma unroll
            for (int c = 0; c < NCMAX; ++c) acc[c] = fma(kv[u], Ws[(r + u) * NCMAX + c], acc[c]);
          }
        }
      }
    }
  }
  if (q >= Q) return;
#pragma unroll
  for (int t = 0; t < P; ++t)
    atomicMax(colmax + (long long)q * P + t, static_cast<unsigned long long>(__double_as_longlong(mx[t])));
  if (W) {
#pragma unroll
    for (int c = 0; c < NCMAX; ++c)
      if (c < nc) part[((long long)blockIdx.y * Qpad + q) * nc + c] = acc[c];
  }
}

__global__ void mean_finalize_kernel(const double* __restrict__ part, int Qpad, int nsplit, int Q, int nc,
                                     const double* __restrict__ Ct, double* __restrict__ Mk) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Q * nc) return;
  const int q = static_cast<int>(idx / nc), c = static_cast<int>(idx % nc);
  double s = Ct[c];
  for (int sp = 0; sp < nsplit; ++sp) s += part[((long long)sp * Qpad + q) * nc + c];
  Mk[idx] = s;
}

// blob of column tile J, K step ks:  [slice 7][column group 8][k chunk 2][column 8][16 bytes]; column = P (q % QT) + t.
// One thread per frakB COLUMN and 16 consecutive rows i: eight neighbouring threads store 128 contiguous bytes per slice.
template <int P, int SD>
__global__ void __launch_bounds__(128) split_frakb_kernel(const double* __restrict__ Kstar, int ldks,
                                                          const double* __restrict__ G, int Npad, int Q, int nJ,
                                                          const unsigned long long* __restrict__ colmax,
                                                          int8_t* __restrict__ blob, double* __restrict__ colscale) {
  constexpr int QT = TN / P;
  constexpr int S = SD, B_STEP = SD * TN * KSTEP;
  const int cg = blockIdx.x * blockDim.x + threadIdx.x;  // global column
  if (cg >= nJ * TN) return;
  const int i0 = blockIdx.y * 16;
  const int J = cg / TN, c = cg % TN;
  const int qq = c / P, t = c % P;
  const int q = J * QT + qq;
  const bool live = qq < QT && q < Q;  // trailing columns of a tile (64 - P QT) and queries past Q: zero digits, zero scale
  const double sc = live ? scale_of(__longlong_as_double(static_cast<long long>(colmax[(long long)q * P + t]))) : 0.0;
  const double inv = live ? 1.0 / sc : 0.0;
  if (blockIdx.y == 0) colscale[cg] = sc;
  uint32_t w[S][4];
#pragma unroll
  for (int s = 0; s < S; ++s) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
  if (live) {
    double kv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) kv[k] = Kstar[(long long)(i0 + k) * ldks + q];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      int d[S];
      digits_of(kv[k] * __ldg(G + (long long)(i0 + k) * P + t) * inv, d);
#pragma unroll
      for (int s = 0; s < S; ++s) w[s][k / 4] |= static_cast<uint32_t>(d[s] & 0xFF) << (8 * (k % 4));
    }
  }
  int8_t* dst = blob + ((long long)J * (Npad / KSTEP) + i0 / KSTEP) * B_STEP + ((i0 / 16) % 2) * 128 + (c / 8) * 256 +
                (c % 8) * 16;
#pragma unroll
  for (int s = 0; s < S; ++s)
    *reinterpret_cast<uint4*>(dst + s * (TN * KSTEP)) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
}

// ---- the contraction --------------------------------------------------------------------------------------------------
struct VarArgs {
  const int8_t* Ablob;
  const int8_t* Bblob;
  const double* rowscale;
  const double* colscale;
  double* Spart;  // [nb][Qpad][NP]
  int nb;         // row blocks of L^-1 (Npad / 128)
  int nJg;        // groups of CL column tiles (Qpad / (CL QT))
  int Qpad;
  long long total_tiles;  // work items: group ceil(nb / group) nJg
  int group;              // row blocks per scheduling group (see tile_of)
  int skip;               // development aid (bcbf_oz_debug_skip_loads): bit 0 / 1 = do not copy the A / B digits (results void)
  unsigned long long* dbg;
};

// tile order: groups of 4 row blocks (longest K extent first) x all column-tile groups; consecutive work items = 4 row
// blocks x consecutive column tiles, so CTAs running side by side share L^-1 digits ~37-fold and frakB digits 4-fold
// through L2.  A work item is (row block I, group of CL column tiles); CTA `rank` of the cluster takes tile CL Jg + rank.
__device__ __forceinline__ bool tile_of(const VarArgs& a, long long t, int& I, int& Jg) {
  const int G = a.group;
  const int per_group = G * a.nJg;
  const int g = static_cast<int>(t / per_group), r = static_cast<int>(t % per_group);
  Jg = r / G;
  const int ii = ((r % G) + static_cast<int>((t / G) % G)) % G;
  I = a.nb - 1 - (G * g + ii);
  return I >= 0;
}

template <int SD = S>
__device__ __forceinline__ void issue_kstep(uint32_t tmem, uint32_t a_base, uint32_t b_base, bool first) {
  constexpr int S = SD;
#pragma unroll
  for (int a = 0; a < S; ++a) {
    const uint64_t ad = smem_desc_kmajor(a_base + a * (TM * KSTEP), 128, 256);
    int b = 0;
    while (b <= S - 1 - a) {
      int nb = S - a - b;
      if (nb > 4) nb = 4;
      const uint64_t bd = smem_desc_kmajor(b_base + b * (TN * KSTEP), 128, 256);
      mma_s8(tmem + (a + b) * TN, ad, bd, idesc_s8(TM, TN * nb), (first && a == 0) ? 0u : 1u);
      b += nb;
    }
  }
}

// CL = 1 (default): every CTA works alone.  CL = 2: clusters of two CTAs work on the same row block and neighbouring
// column tiles and share the L^-1 digits: each CTA fetches half of the 28 KB blob and multicasts it into both shared
// memories (L2 -> SM requests per K step 42 KB -> 28 KB per CTA).  Measured: bit-identical results, NO speed-up (1070 vs
// 1050 cycles per K step with counters on) - the limiter is the shared-memory port, which carries the MMA operand reads
// (96 KB per K step) AND the incoming copies (42 KB): 138 KB / 128 B/clk = 1080 cycles against the 920-cycle MMA pattern,
// and multicast does not change what arrives in shared memory.  Kept as an option (bcbf_oz_set_cluster).
// SD = digits per operand: 7 (default; products with digit sum <= 6: 28 MMAs' worth per K step, 2^-56 truncation) or 6
// (digit sum <= 5: 21 products, 2^-48 truncation: B_k to ~3e-11 of the prior scale, 25 % less tensor work; opt-in).
template <int P, int CL, int SD>
__global__ void __launch_bounds__(kThreads, 1) oz_var_kernel(VarArgs a) {
  constexpr int QT = TN / P, NP = P * (P + 1) / 2;
  constexpr int S = SD, A_STEP = SD * TM * KSTEP, B_STEP = SD * TN * KSTEP, STAGE = A_STEP + B_STEP;
  const int rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const long long worker = blockIdx.x / CL, nworkers = gridDim.x / CL;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tmem_full = empty + NSTAGE;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  double* cs = reinterpret_cast<double*>(smem + NSTAGE * STAGE + 128);
  double* red = cs + TN;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CL);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kEpiThreads);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peer's barriers are initialised before anything is sent to them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {  // ===== producer: bulk copies of the digit blobs =====
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      for (long long t = worker; t < a.total_tiles; t += nworkers) {
        int I, J;
        if (!tile_of(a, t, I, J)) continue;
        J = J * CL + rank;
        const int nks = 4 * (I + 1);
        const int8_t* ap = a.Ablob + 2LL * I * (I + 1) * A_STEP;
        const int8_t* bp = a.Bblob + (long long)J * (a.nb * 4) * B_STEP;
        for (int ks = 0; ks < nks; ++ks) {
          if (a.dbg) {
            const long long c0 = clock64();
            mbar_wait(&empty[stage], phase ^ 1u);
            w_empty += clock64() - c0;
          } else {
            mbar_wait(&empty[stage], phase ^ 1u);
          }
          uint8_t* dst = smem + stage * STAGE;
          if (CL == 1 && (a.skip & 3) != 0) {   // timing experiment: how much of a K step is the incoming copies' share of the port
            const uint32_t tx = ((a.skip & 1) ? 0 : A_STEP) + ((a.skip & 2) ? 0 : B_STEP);
            if (tx == 0) mbar_arrive(&full[stage]);
            else mbar_arrive_expect_tx(&full[stage], tx);
            if (!(a.skip & 1)) bulk_g2s(dst, ap + (long long)ks * A_STEP, A_STEP, &full[stage]);
            if (!(a.skip & 2)) bulk_g2s(dst + A_STEP, bp + (long long)ks * B_STEP, B_STEP, &full[stage]);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
            continue;
          }
          mbar_arrive_expect_tx(&full[stage], STAGE);
          if (CL == 1) {
            bulk_g2s(dst, ap + (long long)ks * A_STEP, A_STEP, &full[stage]);
          } else {  // my 1/CL of the L^-1 digits goes to every CTA of the cluster; the peers send the rest
            constexpr int PART = A_STEP / CL;
            static_assert(PART % 16 == 0, "bulk copies move multiples of 16 bytes");
            bulk_g2s_multicast(dst + rank * PART, ap + (long long)ks * A_STEP + rank * PART, PART, &full[stage],
                               (1u << CL) - 1u);
          }
          bulk_g2s(dst + A_STEP, bp + (long long)ks * B_STEP, B_STEP, &full[stage]);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
      if (CL > 1) {  // tail: every arrival the peer still owes my barriers has landed before this CTA may exit
        for (int s = 0; s < NSTAGE; ++s) {
          mbar_wait(&empty[stage], phase ^ 1u);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
      if (a.dbg) atomicAdd(a.dbg + 3, static_cast<unsigned long long>(w_empty));
    }
  } else if (warp == 5) {
    if (lane == 0) {  // ===== MMA issuer =====
      int stage = 0;
      uint32_t phase = 0, tile_iter = 0;
      long long w_full = 0, w_tmem = 0, nstep = 0;
      const long long t_begin = clock64();
      for (long long t = worker; t < a.total_tiles; t += nworkers) {
        int I, J;
        if (!tile_of(a, t, I, J)) continue;
        J = J * CL + rank;
        const int nks = 4 * (I + 1);
        if (a.dbg) {
          const long long c0 = clock64();
          mbar_wait(tmem_empty, (tile_iter & 1u) ^ 1u);
          w_tmem += clock64() - c0;
          nstep += nks;
        } else {
          mbar_wait(tmem_empty, (tile_iter & 1u) ^ 1u);  // epilogue has drained the accumulators of the previous tile
        }
        tc_fence_after();
        for (int ks = 0; ks < nks; ++ks) {
          if (a.dbg) {
            const long long c0 = clock64();
            mbar_wait(&full[stage], phase);
            w_full += clock64() - c0;
          } else {
            mbar_wait(&full[stage], phase);
          }
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE);
          issue_kstep<SD>(tmem, sa, sa + A_STEP, ks == 0);
          if (CL == 1) mma_commit(&empty[stage]);  // frees the stage when these MMAs have read it
          else mma_commit_multicast(&empty[stage], (1u << CL) - 1u);  // ... in every CTA: the peers write into my stage
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
        mma_commit(tmem_full);
        ++tile_iter;
      }
      if (a.dbg) {  // [0] issue-thread cycles, [1] of which waiting for operands, [2] for the epilogue, [4] K steps
        atomicAdd(a.dbg + 0, static_cast<unsigned long long>(clock64() - t_begin));
        atomicAdd(a.dbg + 1, static_cast<unsigned long long>(w_full));
        atomicAdd(a.dbg + 2, static_cast<unsigned long long>(w_tmem));
        atomicAdd(a.dbg + 4, static_cast<unsigned long long>(nstep));
      }
    }
  } else {  // ===== epilogue warps 0..3: TMEM lanes 32 warp .. 32 warp + 31 =====
    const int tid = threadIdx.x;
    uint32_t tile_iter = 0;
    for (long long t = worker; t < a.total_tiles; t += nworkers) {
      int I, J;
      if (!tile_of(a, t, I, J)) continue;
      J = J * CL + rank;
      if (tid < TN) cs[tid] = a.colscale[(long long)J * TN + tid];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(tmem_full, tile_iter & 1u);
      tc_fence_after();
      if (a.skip & 8) {   // timing experiment: no drain at all
        tc_fence_before();
        mbar_arrive(tmem_empty);
        ++tile_iter;
        continue;
      }
      double V[TN];
#pragma unroll
      for (int c = 0; c < TN; ++c) V[c] = 0.0;
      const uint32_t tbase = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
      for (int d = S - 1; d >= 0; --d) {
        const double w = __longlong_as_double((1023LL - 8 * (d + 2)) << 52);  // 256^-(d+2)
#pragma unroll
        for (int c4 = 0; c4 < TN / 16; ++c4) {
          uint32_t r[16];
          tmem_ld16(tbase + d * TN + c4 * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) V[c4 * 16 + j] = fma(static_cast<double>(static_cast<int>(r[j])), w, V[c4 * 16 + j]);
        }
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);  // accumulators are in registers: the next tile's MMAs may start
      if (a.skip & 4) {   // timing experiment: drain only, no FP64 recombination / Gram products
        if (V[0] == 1.2345e-300) a.Spart[0] = V[1];   // keeps the loads alive
        ++tile_iter;
        continue;
      }
      const double rs = a.rowscale[(long long)I * TM + warp * 32 + lane];
#pragma unroll
      for (int c = 0; c < TN; ++c) V[c] *= rs * cs[c];
#pragma unroll
      for (int q = 0; q < QT; ++q) {
        int e = 0;
#pragma unroll
        for (int x = 0; x < P; ++x)
#pragma unroll
          for (int y = x; y < P; ++y) {
            const double s = warp_sum(V[P * q + x] * V[P * q + y]);
            if (lane == 0) red[warp * (QT * NP) + q * NP + e] = s;
            ++e;
          }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int x = tid; x < QT * NP; x += kEpiThreads) {
        const double s = (red[x] + red[QT * NP + x]) + (red[2 * QT * NP + x] + red[3 * QT * NP + x]);
        const int q = x / NP, e = x % NP;
        a.Spart[((long long)I * a.Qpad + (long long)J * QT + q) * NP + e] = s;
      }
      ++tile_iter;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();
  if (warp == 5) tmem_dealloc(tmem, 512);
}

// Bk[q] = kss B - sum over row blocks of Spart
__global__ void finalize_kernel(const double* __restrict__ Spart, int Qpad, int nb, int Q, int p,
                                const double* __restrict__ Bmat, double kss, double* __restrict__ Bk) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int npair = p * (p + 1) / 2;
  int e = 0;
  for (int i = 0; i < p; ++i)
    for (int j = i; j < p; ++j) {
      double s = 0.0;
      for (int sp = 0; sp < nb; ++sp) s += Spart[((long long)sp * Qpad + q) * npair + e];
      const double v = __dmul_rn(kss, Bmat[i * p + j]) - s;  // no FMA contraction: two roundings, reproducible on the host
      Bk[((long long)q * p + i) * p + j] = v;
      Bk[((long long)q * p + j) * p + i] = v;
      ++e;
    }
}

struct Ws {
  void* ptr = nullptr;
  size_t bytes = 0;
};
static Ws g_ws[12][64];  // 0: frakB digits, 1: colmax, 2: colscale, 3: Spart, 4: mean partials; GEMM: 5: A digits,
                         // 6: B digits, 7: row scales, 8: column maxima + scales; update: 9: PA digits, 10: PB digits, 11: scales

static int workspace(int slot, size_t bytes, void** out) {
  int dev = 0;
  BCBF_CUDA(cudaGetDevice(&dev));
  Ws& w = g_ws[slot][dev & 63];
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

struct Prof {
  bool on = false;
  cudaEvent_t e0[256], e1[256];
  int n = 0;
};
static Prof g_prof;

static unsigned long long* g_dbg = nullptr;

static int g_cluster = 1;  // CTAs per cluster of oz_var_kernel (1, 2 or 4); bcbf_oz_set_cluster
static int g_group = 4;    // row blocks per scheduling group of oz_var_kernel; bcbf_oz_set_group
static int g_skip = 0;     // bcbf_oz_debug_skip_loads

template <int P, int CL, int SD>
static int launch_var(VarArgs a, cudaStream_t stream) {
  constexpr int kSmemBytes = NSTAGE * SD * (TM + TN) * KSTEP + 128 + TN * 8 + 4 * 160 * 8;
  int dev = 0, sms = 148;
  BCBF_CUDA(cudaGetDevice(&dev));
  BCBF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  BCBF_CUDA(cudaFuncSetAttribute(oz_var_kernel<P, CL, SD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int workers = sms / CL;
  if (CL > 1) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // a persistent grid must be co-resident: ask how many clusters fit (SM pairs of one TPC)
    static int max_clusters[64] = {0};
    if (max_clusters[dev & 63] == 0) {
      cfg.gridDim = dim3(sms - sms % CL);
      int n = 0;
      BCBF_CUDA(cudaOccupancyMaxActiveClusters(&n, oz_var_kernel<P, CL, SD>, &cfg));
      max_clusters[dev & 63] = n > 0 ? n : 1;
    }
    if (workers > max_clusters[dev & 63]) workers = max_clusters[dev & 63];
  }
  if (a.total_tiles < workers) workers = static_cast<int>(a.total_tiles);
  cfg.gridDim = dim3(workers * CL);
  const bool prof = g_prof.on && g_prof.n < 256;
  if (prof) {
    BCBF_CUDA(cudaEventCreate(&g_prof.e0[g_prof.n]));
    BCBF_CUDA(cudaEventCreate(&g_prof.e1[g_prof.n]));
    BCBF_CUDA(cudaEventRecord(g_prof.e0[g_prof.n], stream));
  }
  BCBF_CUDA(cudaLaunchKernelEx(&cfg, oz_var_kernel<P, CL, SD>, a));
  BCBF_LAUNCH_CHECK();
  if (prof) {
    BCBF_CUDA(cudaEventRecord(g_prof.e1[g_prof.n], stream));
    ++g_prof.n;
  }
  return BCBF_OK;
}

template <int P, int SD>
static int run_blocks(const int8_t* Ablob, const double* rowscale, int Npad, const double* Kstar, int ldks,
                      const double* G, const double* W, const double* Bmat, const double* Ct, double kss, int n, int Q,
                      double* Mk, double* Bk, cudaStream_t stream) {
  constexpr int QT = TN / P, NP = P * (P + 1) / 2;
  constexpr int B_STEP = SD * TN * KSTEP;
  const int CL = g_cluster;
  const int nb = Npad / TM, nJg = ceil_div(Q, QT * CL), nJ = nJg * CL, Qpad = nJ * QT, nc = n * P;
  const int nsplit = ceil_div(Npad, kCmRows);
  void *bblob, *colmax, *colscale, *spart, *mpart = nullptr;
  int rc;
  if ((rc = workspace(1, sizeof(unsigned long long) * (size_t)Qpad * P, &colmax))) return rc;
  if (Mk && (rc = workspace(4, sizeof(double) * (size_t)nsplit * Qpad * nc, &mpart))) return rc;
  BCBF_CUDA(cudaMemsetAsync(colmax, 0, sizeof(unsigned long long) * (size_t)Qpad * P, stream));
  {
    const dim3 grid(ceil_div(Q, 128), nsplit);
    unsigned long long* cm = static_cast<unsigned long long*>(colmax);
    double* mp = static_cast<double*>(mpart);
    const double* Wm = Mk ? W : nullptr;
    if (nc <= 4) colmax_mean_kernel<P, 4><<<grid, 128, 0, stream>>>(Kstar, ldks, G, Wm, nc, Npad, Q, Qpad, cm, mp);
    else if (nc <= 9) colmax_mean_kernel<P, 9><<<grid, 128, 0, stream>>>(Kstar, ldks, G, Wm, nc, Npad, Q, Qpad, cm, mp);
    else if (nc <= 16) colmax_mean_kernel<P, 16><<<grid, 128, 0, stream>>>(Kstar, ldks, G, Wm, nc, Npad, Q, Qpad, cm, mp);
    else colmax_mean_kernel<P, 32><<<grid, 128, 0, stream>>>(Kstar, ldks, G, Wm, nc, Npad, Q, Qpad, cm, mp);
  }
  BCBF_LAUNCH_CHECK();
  if (Mk) {
    mean_finalize_kernel<<<ceil_div((long long)Q * nc, 256), 256, 0, stream>>>(static_cast<const double*>(mpart), Qpad,
                                                                               nsplit, Q, nc, Ct, Mk);
    BCBF_LAUNCH_CHECK();
  }
  if (!Bk) return BCBF_OK;
  if ((rc = workspace(0, (size_t)nJ * (Npad / KSTEP) * B_STEP, &bblob))) return rc;
  if ((rc = workspace(2, sizeof(double) * (size_t)nJ * TN, &colscale))) return rc;
  if ((rc = workspace(3, sizeof(double) * (size_t)nb * Qpad * NP, &spart))) return rc;
  // every column of every tile is written (dead columns and queries past Q: zero digits, zero scale)
  split_frakb_kernel<P, SD><<<dim3(ceil_div(nJ * TN, 128), Npad / 16), 128, 0, stream>>>(
      Kstar, ldks, G, Npad, Q, nJ, static_cast<const unsigned long long*>(colmax), static_cast<int8_t*>(bblob),
      static_cast<double*>(colscale));
  BCBF_LAUNCH_CHECK();
  VarArgs a{};
  a.Ablob = Ablob;
  a.Bblob = static_cast<const int8_t*>(bblob);
  a.rowscale = rowscale;
  a.colscale = static_cast<const double*>(colscale);
  a.Spart = static_cast<double*>(spart);
  a.nb = nb;
  a.nJg = nJg;
  a.Qpad = Qpad;
  a.group = g_group;
  a.skip = g_skip;
  a.total_tiles = (long long)ceil_div(nb, g_group) * g_group * nJg;
  a.dbg = g_dbg;
  rc = CL == 4   ? launch_var<P, 4, SD>(a, stream)
       : CL == 2 ? launch_var<P, 2, SD>(a, stream)
                 : launch_var<P, 1, SD>(a, stream);
  if (rc) return rc;
  finalize_kernel<<<ceil_div(Q, 128), 128, 0, stream>>>(a.Spart, Qpad, nb, Q, P, Bmat, kss, Bk);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

// ======================================================================================================================
// General FP64-accurate GEMM on the int8 tensor cores:  C (M,N) = alpha A (M,K) B (K,N), row-major, with optional
// triangular structure of one operand.  Same digit arithmetic, MMA pattern and pipeline as oz_var_kernel; the epilogue
// stores the recombined FP64 tile.  Used by bcbf_trtri for the large levels of the triangular inverse
// (T = L21 X11, X21 = -X22 T: control_affine_model.py:565's solves become products with L^-1).
// ======================================================================================================================
constexpr int kTriGemmNone = 0, kTriGemmALower = 1, kTriGemmBLower = 2;
constexpr int kTriGemmTnLower = 3;  // C = A^T B with A and B lower triangular: term k contributes only for k >= max(i, j)

// Inner-dimension balancing: C = (A D)(D^-1 B) for any diagonal D.  With one scale per row of A and per column of B, an
// operand whose magnitude falls steeply along k (rows of a Cholesky factor) loses the bits of its small entries even
// though they meet large entries of the other operand (rows of its inverse).  D_k = 2^round(log2 sqrt(rowmax_k(B) /
// colmax_k(A))) equalises the two profiles; powers of two, so the products are unchanged.
__global__ void colmax_of_a_kernel(const double* __restrict__ A, int lda, int M, int K, int tri, int rows_per_block,
                                   unsigned long long* __restrict__ amax) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  int i0 = blockIdx.y * rows_per_block;
  const int i1 = min(M, i0 + rows_per_block);
  if (tri == kTriGemmALower) i0 = max(i0, k);
  double mx = 0.0;
  int i = i0;
  for (; i + 8 <= i1; i += 8) {
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = A[(long long)(i + u) * lda + k];
#pragma unroll
    for (int u = 0; u < 8; ++u) mx = fmax(mx, fabs(v[u]));
  }
  for (; i < i1; ++i) mx = fmax(mx, fabs(A[(long long)i * lda + k]));
  atomicMax(amax + k, static_cast<unsigned long long>(__double_as_longlong(mx)));
}

__global__ void kscale_kernel(const double* __restrict__ B, int ldb, int K, int N, int tri,
                              const unsigned long long* __restrict__ amax, double* __restrict__ kscale) {
  const int k = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (k >= K) return;
  const int jend = tri == kTriGemmBLower ? min(N, k + 1) : N;
  double mx = 0.0;
  for (int j = lane; j < jend; j += 32) mx = fmax(mx, fabs(B[(long long)k * ldb + j]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) {
    const double ca = __longlong_as_double(static_cast<long long>(amax[k]));
    double d = 1.0;
    if (mx > 0.0 && ca > 0.0) {
      int eb, ea;
      frexp(mx, &eb);
      frexp(ca, &ea);
      int e = eb - ea;               // log2(rowmax_B / colmax_A), rounded; D = 2^(e/2)
      e = (e >= 0 ? e + 1 : e) / 2;
      d = ldexp(1.0, e);
    }
    kscale[k] = d;
  }
}

__global__ void rowscale_rect_kernel(const double* __restrict__ A, int lda, int M, int K, int tri,
                                     const double* __restrict__ kscale, double* __restrict__ rowscale) {
  const int row = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, lane = threadIdx.x % 32;
  if (row >= M) return;
  const int kend = tri == kTriGemmALower ? min(K, row + 1) : K;
  double mx = 0.0;
  for (int k = lane; k < kend; k += 32) mx = fmax(mx, fabs(A[(long long)row * lda + k]) * (kscale ? kscale[k] : 1.0));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) rowscale[row] = scale_of(mx);
}

// A D (M,K) row-major -> blobs [(I nks + ks)] of [digit][row group ROWS/8][k chunk 2][row 8][16 B]; entries with k > row
// are taken as zero when A is lower triangular (the strictly upper part of the storage is not read).  ROWS = 128: the
// MMA's M-side operand; ROWS = 64: the N-side operand of a product with A^T (rank-k updates C += alpha P P^T).
template <int ROWS>
__global__ void __launch_bounds__(2 * ROWS) split_rows_kernel(const double* __restrict__ A, int lda, int K, int tri,
                                                              const double* __restrict__ kscale,
                                                              const double* __restrict__ rowscale,
                                                              int8_t* __restrict__ blob) {
  constexpr int STEP = S * ROWS * KSTEP;
  const int kc = blockIdx.x, I = blockIdx.y;  // 128-wide k block, ROWS-row block
  const int nks = K / KSTEP;
  const int r = threadIdx.x % ROWS, half = threadIdx.x / ROWS;
  const long long row = (long long)I * ROWS + r;
  const double inv = 1.0 / rowscale[row];
  int8_t* base = blob + ((long long)I * nks + 4LL * kc) * STEP;
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    const int c16 = half * 4 + c;
    const int k0 = kc * 128 + c16 * 16;
    if (k0 >= K) break;
    const double* src = A + row * lda + k0;
    uint32_t w[S][4];
#pragma unroll
    for (int s = 0; s < S; ++s) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const double v =
          (tri == kTriGemmALower && k0 + k > row) ? 0.0 : src[k] * (kscale ? __ldg(kscale + k0 + k) : 1.0);
      int d[S];
      digits_of(v * inv, d);
#pragma unroll
      for (int s = 0; s < S; ++s) w[s][k / 4] |= static_cast<uint32_t>(d[s] & 0xFF) << (8 * (k % 4));
    }
    int8_t* dst = base + (long long)(c16 / 2) * STEP + (r / 8) * 256 + (c16 % 2) * 128 + (r % 8) * 16;
#pragma unroll
    for (int s = 0; s < S; ++s)
      *reinterpret_cast<uint4*>(dst + s * (ROWS * KSTEP)) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
  }
}

// column maxima of D^-1 B, B (K,N) row-major (entries with k < j are taken as zero when B is lower triangular)
__global__ void colmax_rect_kernel(const double* __restrict__ B, int ldb, int K, int N, int tri, int rows_per_block,
                                   const double* __restrict__ kscale, unsigned long long* __restrict__ colmax) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  int k0 = blockIdx.y * rows_per_block;
  const int k1 = min(K, k0 + rows_per_block);
  if (tri == kTriGemmBLower) k0 = max(k0, j);
  double mx = 0.0;
  int k = k0;
  for (; k + 8 <= k1; k += 8) {  // eight independent loads in flight
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = B[(long long)(k + u) * ldb + j];
#pragma unroll
    for (int u = 0; u < 8; ++u) mx = fmax(mx, kscale ? fabs(v[u]) / __ldg(kscale + k + u) : fabs(v[u]));
  }
  for (; k < k1; ++k) mx = fmax(mx, kscale ? fabs(B[(long long)k * ldb + j]) / __ldg(kscale + k) : fabs(B[(long long)k * ldb + j]));
  atomicMax(colmax + j, static_cast<unsigned long long>(__double_as_longlong(mx)));
}

// D^-1 B (K,N) row-major -> blobs [(J nks + ks)] of [digit][column group COLS/8][k chunk 2][column 8][16 B], one thread
// per column and 16 consecutive k.  COLS = 64: the MMA's N-side operand; COLS = 128: the M-side operand of a product with
// B^T on the left (C = A^T B).
template <int COLS>
__global__ void __launch_bounds__(128) split_cols_kernel(const double* __restrict__ B, int ldb, int K, int N, int tri,
                                                         const double* __restrict__ kscale,
                                                         const unsigned long long* __restrict__ colmax,
                                                         int8_t* __restrict__ blob, double* __restrict__ colscale) {
  constexpr int STEP = S * COLS * KSTEP;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const int k0 = blockIdx.y * 16;
  const int J = j / COLS, c = j % COLS;
  const double sc = scale_of(__longlong_as_double(static_cast<long long>(colmax[j])));
  const double inv = 1.0 / sc;
  if (blockIdx.y == 0) colscale[j] = sc;
  uint32_t w[S][4];
#pragma unroll
  for (int s = 0; s < S; ++s) w[s][0] = w[s][1] = w[s][2] = w[s][3] = 0u;
  if (!(tri == kTriGemmBLower && k0 + 15 < j)) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double v = (tri == kTriGemmBLower && k0 + k < j) ? 0.0 : B[(long long)(k0 + k) * ldb + j];
      if (kscale) v /= __ldg(kscale + k0 + k);
      int d[S];
      digits_of(v * inv, d);
#pragma unroll
      for (int s = 0; s < S; ++s) w[s][k / 4] |= static_cast<uint32_t>(d[s] & 0xFF) << (8 * (k % 4));
    }
  }
  int8_t* dst = blob + ((long long)J * (K / KSTEP) + k0 / KSTEP) * STEP + ((k0 / 16) % 2) * 128 + (c / 8) * 256 +
                (c % 8) * 16;
#pragma unroll
  for (int s = 0; s < S; ++s)
    *reinterpret_cast<uint4*>(dst + s * (COLS * KSTEP)) = make_uint4(w[s][0], w[s][1], w[s][2], w[s][3]);
}

struct GemmI8Args {
  const int8_t* Ablob;
  const int8_t* Bblob;
  const double* rowscale;
  const double* colscale;
  double* C;
  long long ldc;
  double alpha;
  int nI, nJ, nks, tri;
  long long total_tiles;
};

// tile t -> (I, J) and its K-step range.  Order: row-block groups of 4 x all column tiles (operand sharing through L2 as
// in oz_var_kernel); with a lower-triangular A the longest row blocks come first.
__device__ __forceinline__ bool gemm_tile_of(const GemmI8Args& a, long long t, int& I, int& J, int& ks0, int& ks1) {
  const int per_group = 4 * a.nJ;
  const int g = static_cast<int>(t / per_group), r = static_cast<int>(t % per_group);
  J = r >> 2;
  const int ii = ((r & 3) + static_cast<int>((t >> 2) & 3)) & 3;
  I = a.nI - 1 - (4 * g + ii);
  if (I < 0) return false;
  ks0 = a.tri == kTriGemmBLower ? (J * TN) / KSTEP : (a.tri == kTriGemmTnLower ? max(I * TM, J * TN) / KSTEP : 0);
  ks1 = a.tri == kTriGemmALower ? min(a.nks, 4 * (I + 1)) : a.nks;
  return ks1 > ks0;
}

__global__ void __launch_bounds__(kThreads, 1) oz_gemm_kernel(GemmI8Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tmem_full = empty + NSTAGE;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);
  double* cs = reinterpret_cast<double*>(smem + NSTAGE * STAGE + 128);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kEpiThreads);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
        int I, J, ks0, ks1;
        if (!gemm_tile_of(a, t, I, J, ks0, ks1)) continue;
        const int8_t* ap = a.Ablob + (long long)I * a.nks * A_STEP;
        const int8_t* bp = a.Bblob + (long long)J * a.nks * B_STEP;
        for (int ks = ks0; ks < ks1; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* dst = smem + stage * STAGE;
          mbar_arrive_expect_tx(&full[stage], STAGE);
          bulk_g2s(dst, ap + (long long)ks * A_STEP, A_STEP, &full[stage]);
          bulk_g2s(dst + A_STEP, bp + (long long)ks * B_STEP, B_STEP, &full[stage]);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, tile_iter = 0;
      for (long long t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
        int I, J, ks0, ks1;
        if (!gemm_tile_of(a, t, I, J, ks0, ks1)) continue;
        mbar_wait(tmem_empty, (tile_iter & 1u) ^ 1u);
        tc_fence_after();
        for (int ks = ks0; ks < ks1; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE);
          issue_kstep(tmem, sa, sa + A_STEP, ks == ks0);
          mma_commit(&empty[stage]);
          if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
        }
        mma_commit(tmem_full);
        ++tile_iter;
      }
    }
  } else {
    const int tid = threadIdx.x;
    uint32_t tile_iter = 0;
    for (long long t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
      int I, J, ks0, ks1;
      if (!gemm_tile_of(a, t, I, J, ks0, ks1)) continue;
      asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone is done with the previous tile's column scales
      if (tid < TN) cs[tid] = a.colscale[(long long)J * TN + tid];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(tmem_full, tile_iter & 1u);
      tc_fence_after();
      const uint32_t tbase = tmem + (static_cast<uint32_t>(warp * 32) << 16);
      const long long row = (long long)I * TM + warp * 32 + lane;
      const double rs = a.alpha * a.rowscale[row];
      double* crow = a.C + row * a.ldc + (long long)J * TN;
#pragma unroll
      for (int c4 = 0; c4 < TN / 16; ++c4) {
        double v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.0;
#pragma unroll
        for (int d = S - 1; d >= 0; --d) {
          const double w = __longlong_as_double((1023LL - 8 * (d + 2)) << 52);
          uint32_t r[16];
          tmem_ld16(tbase + d * TN + c4 * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fma(static_cast<double>(static_cast<int>(r[j])), w, v[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; j += 2)
          *reinterpret_cast<double2*>(crow + c4 * 16 + j) =
              make_double2(v[j] * (rs * cs[c4 * 16 + j]), v[j + 1] * (rs * cs[c4 * 16 + j + 1]));
      }
      tc_fence_before();
      mbar_arrive(tmem_empty);
      ++tile_iter;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 512);
}

static int run_gemm(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C,
                    int ldc, int tri, cudaStream_t stream) {
  const int nI = M / TM, nJ = N / TN, nks = K / KSTEP;
  void *ablob, *bblob, *rows, *cols;
  int rc;
  if ((rc = workspace(5, (size_t)nI * nks * A_STEP, &ablob))) return rc;
  if ((rc = workspace(6, (size_t)nJ * nks * B_STEP, &bblob))) return rc;
  if ((rc = workspace(7, sizeof(double) * (size_t)M, &rows))) return rc;
  if ((rc = workspace(8, 16 * ((size_t)N + (size_t)K), &cols))) return rc;
  // workspace 8: [colmax N | colscale N | amax K | kscale K]
  unsigned long long* colmax = static_cast<unsigned long long*>(cols);
  double* colscale = reinterpret_cast<double*>(colmax + N);
  unsigned long long* amax = reinterpret_cast<unsigned long long*>(colscale + N);
  double* kscale = reinterpret_cast<double*>(amax + K);
  BCBF_CUDA(cudaMemsetAsync(amax, 0, sizeof(unsigned long long) * (size_t)K, stream));
  colmax_of_a_kernel<<<dim3(ceil_div(K, 128), ceil_div(M, 256)), 128, 0, stream>>>(A, lda, M, K, tri, 256, amax);
  BCBF_LAUNCH_CHECK();
  kscale_kernel<<<ceil_div(K, 8), 256, 0, stream>>>(B, ldb, K, N, tri, amax, kscale);
  BCBF_LAUNCH_CHECK();
  rowscale_rect_kernel<<<ceil_div(M, 8), 256, 0, stream>>>(A, lda, M, K, tri, kscale, static_cast<double*>(rows));
  BCBF_LAUNCH_CHECK();
  split_rows_kernel<TM><<<dim3(ceil_div(K, TM), nI), 256, 0, stream>>>(A, lda, K, tri, kscale,
                                                                  static_cast<const double*>(rows),
                                                                  static_cast<int8_t*>(ablob));
  BCBF_LAUNCH_CHECK();
  BCBF_CUDA(cudaMemsetAsync(colmax, 0, sizeof(unsigned long long) * (size_t)N, stream));
  colmax_rect_kernel<<<dim3(ceil_div(N, 128), ceil_div(K, 256)), 128, 0, stream>>>(B, ldb, K, N, tri, 256, kscale,
                                                                                   colmax);
  BCBF_LAUNCH_CHECK();
  split_cols_kernel<TN><<<dim3(ceil_div(N, 128), K / 16), 128, 0, stream>>>(B, ldb, K, N, tri, kscale, colmax,
                                                                       static_cast<int8_t*>(bblob), colscale);
  BCBF_LAUNCH_CHECK();
  GemmI8Args a{};
  a.Ablob = static_cast<const int8_t*>(ablob);
  a.Bblob = static_cast<const int8_t*>(bblob);
  a.rowscale = static_cast<const double*>(rows);
  a.colscale = colscale;
  a.C = C;
  a.ldc = ldc;
  a.alpha = alpha;
  a.nI = nI;
  a.nJ = nJ;
  a.nks = nks;
  a.tri = tri;
  a.total_tiles = (long long)ceil_div(nI, 4) * 4 * nJ;
  int dev = 0, sms = 148;
  BCBF_CUDA(cudaGetDevice(&dev));
  BCBF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  BCBF_CUDA(cudaFuncSetAttribute(oz_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  const int grid = a.total_tiles < sms ? static_cast<int>(a.total_tiles) : sms;
  oz_gemm_kernel<<<grid, kThreads, kSmemBytes, stream>>>(a);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

// C (M,N) = alpha A^T B with A (K,M) and B (K,N) row-major.  lower != 0: both are square lower triangular (K = M = N), the
// K steps below max(i, j) are skipped (Kb^-1 = L^-T L^-1 of the log-marginal gradient).  Both operands are
// split by columns; no inner balancing (for A = B the two K profiles coincide).
static int run_gemm_tn(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C,
                       int ldc, int lower, cudaStream_t stream) {
  const int nI = M / TM, nJ = N / TN, nks = K / KSTEP;
  const int tri = lower ? kTriGemmBLower : kTriGemmNone;
  void *ablob, *bblob, *cols;
  int rc;
  if ((rc = workspace(5, (size_t)nI * nks * A_STEP, &ablob))) return rc;
  if ((rc = workspace(6, (size_t)nJ * nks * B_STEP, &bblob))) return rc;
  if ((rc = workspace(8, 16 * ((size_t)N + (size_t)M), &cols))) return rc;
  unsigned long long* cmB = static_cast<unsigned long long*>(cols);
  double* csB = reinterpret_cast<double*>(cmB + N);
  unsigned long long* cmA = reinterpret_cast<unsigned long long*>(csB + N);
  double* csA = reinterpret_cast<double*>(cmA + M);
  BCBF_CUDA(cudaMemsetAsync(cmB, 0, sizeof(unsigned long long) * (size_t)N, stream));
  BCBF_CUDA(cudaMemsetAsync(cmA, 0, sizeof(unsigned long long) * (size_t)M, stream));
  colmax_rect_kernel<<<dim3(ceil_div(M, 128), ceil_div(K, 256)), 128, 0, stream>>>(A, lda, K, M, tri, 256, nullptr, cmA);
  BCBF_LAUNCH_CHECK();
  split_cols_kernel<TM><<<dim3(ceil_div(M, 128), K / 16), 128, 0, stream>>>(A, lda, K, M, tri, nullptr, cmA,
                                                                           static_cast<int8_t*>(ablob), csA);
  BCBF_LAUNCH_CHECK();
  colmax_rect_kernel<<<dim3(ceil_div(N, 128), ceil_div(K, 256)), 128, 0, stream>>>(B, ldb, K, N, tri, 256, nullptr, cmB);
  BCBF_LAUNCH_CHECK();
  split_cols_kernel<TN><<<dim3(ceil_div(N, 128), K / 16), 128, 0, stream>>>(B, ldb, K, N, tri, nullptr, cmB,
                                                                           static_cast<int8_t*>(bblob), csB);
  BCBF_LAUNCH_CHECK();
  GemmI8Args a{};
  a.Ablob = static_cast<const int8_t*>(ablob);
  a.Bblob = static_cast<const int8_t*>(bblob);
  a.rowscale = csA;
  a.colscale = csB;
  a.C = C;
  a.ldc = ldc;
  a.alpha = alpha;
  a.nI = nI;
  a.nJ = nJ;
  a.nks = nks;
  a.tri = lower ? kTriGemmTnLower : kTriGemmNone;
  a.total_tiles = (long long)ceil_div(nI, 4) * 4 * nJ;
  int dev = 0, sms = 148;
  BCBF_CUDA(cudaGetDevice(&dev));
  BCBF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  BCBF_CUDA(cudaFuncSetAttribute(oz_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  const int grid = a.total_tiles < sms ? static_cast<int>(a.total_tiles) : sms;
  oz_gemm_kernel<<<grid, kThreads, kSmemBytes, stream>>>(a);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

// ======================================================================================================================
// Rank-K update on the int8 tensor cores:  C (M,N) += alpha PA (M,K) PB (N,K)^T, all tiles or only those that touch the
// lower triangle (the Cholesky trailing update A22 -= L21 L21^T, make_psd's factorisation control_affine_model.py:907).
// One CTA per 128 x 64 tile, NOT persistent: the look-ahead of bcbf_potrf needs SMs to come free every few microseconds
// for the high-priority panel kernels.  Same digit arithmetic / MMA pattern as above; the epilogue is a read-modify-write.
// ======================================================================================================================
struct UpdateArgs {
  const int8_t* Ablob;
  const int8_t* Bblob;
  const double* rowscaleA;
  const double* rowscaleB;
  double* C;
  long long ldc;
  double alpha;
  int nJ, nks, lower;
};

__global__ void __launch_bounds__(kThreads, 1) oz_update_kernel(UpdateArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE);
  uint64_t* empty = full + NSTAGE;
  uint64_t* tmem_full = empty + NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 2);
  double* cs = reinterpret_cast<double*>(smem + NSTAGE * STAGE + 128);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  int I, J;
  if (a.lower) {  // tile t = I (I + 1) + J,  0 <= J < 2 I + 2
    const long long t = blockIdx.x;
    I = static_cast<int>((sqrt(4.0 * static_cast<double>(t) + 1.0) - 1.0) * 0.5);
    while ((long long)I * (I + 1) > t) --I;
    while ((long long)(I + 1) * (I + 2) <= t) ++I;
    J = static_cast<int>(t - (long long)I * (I + 1));
  } else {
    I = blockIdx.x / a.nJ;
    J = blockIdx.x % a.nJ;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 4) {
    if (lane == 0) {
      const int8_t* ap = a.Ablob + (long long)I * a.nks * A_STEP;
      const int8_t* bp = a.Bblob + (long long)J * a.nks * B_STEP;
      int stage = 0;
      uint32_t phase = 0;
      for (int ks = 0; ks < a.nks; ++ks) {
        mbar_wait(&empty[stage], phase ^ 1u);
        uint8_t* dst = smem + stage * STAGE;
        mbar_arrive_expect_tx(&full[stage], STAGE);
        bulk_g2s(dst, ap + (long long)ks * A_STEP, A_STEP, &full[stage]);
        bulk_g2s(dst + A_STEP, bp + (long long)ks * B_STEP, B_STEP, &full[stage]);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ks = 0; ks < a.nks; ++ks) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE);
        issue_kstep(tmem, sa, sa + A_STEP, ks == 0);
        mma_commit(&empty[stage]);
        if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
      }
      mma_commit(tmem_full);
    }
  } else {
    const int tid = threadIdx.x;
    const long long row = (long long)I * TM + warp * 32 + lane;
    double* crow = a.C + row * a.ldc + (long long)J * TN;
#pragma unroll
    for (int l = 0; l < TN * 8 / 128; ++l) asm volatile("prefetch.global.L2 [%0];" ::"l"(crow + l * 16));
    if (tid < TN) cs[tid] = a.rowscaleB[(long long)J * TN + tid];
    const double rs = a.alpha * a.rowscaleA[row];
    asm volatile("bar.sync 1, 128;" ::: "memory");
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const uint32_t tbase = tmem + (static_cast<uint32_t>(warp * 32) << 16);
#pragma unroll
    for (int c4 = 0; c4 < TN / 16; ++c4) {
      double2 cold[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cold[j] = *reinterpret_cast<const double2*>(crow + c4 * 16 + 2 * j);
      double v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.0;
#pragma unroll
      for (int d = S - 1; d >= 0; --d) {
        const double w = __longlong_as_double((1023LL - 8 * (d + 2)) << 52);
        uint32_t r[16];
        tmem_ld16(tbase + d * TN + c4 * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fma(static_cast<double>(static_cast<int>(r[j])), w, v[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<double2*>(crow + c4 * 16 + 2 * j) =
            make_double2(fma(v[2 * j], rs * cs[c4 * 16 + 2 * j], cold[j].x),
                         fma(v[2 * j + 1], rs * cs[c4 * 16 + 2 * j + 1], cold[j].y));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 512);
}

// PA (M,K; lda), PB (N,K; ldb) row-major; PB == PA (same pointer, N <= M) shares the row scales
static int run_update(int M, int N, int K, double alpha, const double* PA, int lda, const double* PB, int ldb, double* C,
                      int ldc, int lower, cudaStream_t stream) {
  const int nI = M / TM, nJ = N / TN, nks = K / KSTEP;
  void *ablob, *bblob, *scales;
  int rc;
  if ((rc = workspace(9, (size_t)nI * nks * A_STEP, &ablob))) return rc;
  if ((rc = workspace(10, (size_t)nJ * nks * B_STEP, &bblob))) return rc;
  if ((rc = workspace(11, sizeof(double) * ((size_t)M + (size_t)N), &scales))) return rc;
  double* rsA = static_cast<double*>(scales);
  double* rsB = rsA + M;
  rowscale_rect_kernel<<<ceil_div(M, 8), 256, 0, stream>>>(PA, lda, M, K, kTriGemmNone, nullptr, rsA);
  BCBF_LAUNCH_CHECK();
  if (PB == PA && ldb == lda) {
    rsB = rsA;
  } else {
    rowscale_rect_kernel<<<ceil_div(N, 8), 256, 0, stream>>>(PB, ldb, N, K, kTriGemmNone, nullptr, rsB);
    BCBF_LAUNCH_CHECK();
  }
  split_rows_kernel<TM><<<dim3(ceil_div(K, 128), nI), 2 * TM, 0, stream>>>(PA, lda, K, kTriGemmNone, nullptr, rsA,
                                                                           static_cast<int8_t*>(ablob));
  BCBF_LAUNCH_CHECK();
  split_rows_kernel<TN><<<dim3(ceil_div(K, 128), nJ), 2 * TN, 0, stream>>>(PB, ldb, K, kTriGemmNone, nullptr, rsB,
                                                                           static_cast<int8_t*>(bblob));
  BCBF_LAUNCH_CHECK();
  UpdateArgs a{};
  a.Ablob = static_cast<const int8_t*>(ablob);
  a.Bblob = static_cast<const int8_t*>(bblob);
  a.rowscaleA = rsA;
  a.rowscaleB = rsB;
  a.C = C;
  a.ldc = ldc;
  a.alpha = alpha;
  a.nJ = nJ;
  a.nks = nks;
  a.lower = lower;
  const long long tiles = lower ? (long long)nI * (nI + 1) : (long long)nI * nJ;
  BCBF_CUDA(cudaFuncSetAttribute(oz_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  oz_update_kernel<<<static_cast<unsigned>(tiles), kThreads, kSmemBytes, stream>>>(a);
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

}  // namespace oz
}  // namespace bcbf

using namespace bcbf;

extern "C" long long bcbf_oz_factor_bytes_d(int Npad, int digits) {
  if (Npad <= 0 || Npad % oz::TM != 0 || (digits != 6 && digits != 7)) return 0;
  const long long nb = Npad / oz::TM;
  return 2LL * nb * (nb + 1) * digits * oz::TM * oz::KSTEP;
}
extern "C" long long bcbf_oz_factor_bytes(int Npad) { return bcbf_oz_factor_bytes_d(Npad, oz::S); }

extern "C" int bcbf_oz_max_npad(void) { return oz::kMaxNpad; }

extern "C" int bcbf_oz_split_factor(const double* Linv, int ld, int Npad, void* digits, double* rowscale,
                                    void* stream_) {
  return bcbf_oz_split_factor_d(Linv, ld, Npad, digits, rowscale, oz::S, stream_);
}

extern "C" int bcbf_oz_split_factor_d(const double* Linv, int ld, int Npad, void* digits, double* rowscale, int ndigits,
                                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(Linv && digits && rowscale, "bcbf_oz_split_factor: null pointer");
  BCBF_REQUIRE(ndigits == 6 || ndigits == 7, "bcbf_oz_split_factor: digits=%d (6 or 7)", ndigits);
  BCBF_REQUIRE(Npad > 0 && Npad % oz::TM == 0 && ld >= Npad && Npad <= oz::kMaxNpad,
               "bcbf_oz_split_factor: Npad=%d ld=%d (Npad must be a multiple of 128 and <= %d)", Npad, ld, oz::kMaxNpad);
  oz::rowscale_kernel<<<ceil_div(Npad, 8), 256, 0, stream>>>(Linv, ld, Npad, rowscale);
  BCBF_LAUNCH_CHECK();
  const int nb = Npad / oz::TM;
  if (ndigits == 7) oz::split_factor_kernel<7><<<dim3(nb, nb), 256, 0, stream>>>(Linv, ld, rowscale, static_cast<int8_t*>(digits));
  else oz::split_factor_kernel<6><<<dim3(nb, nb), 256, 0, stream>>>(Linv, ld, rowscale, static_cast<int8_t*>(digits));
  BCBF_LAUNCH_CHECK();
  return BCBF_OK;
}

extern "C" int bcbf_posterior_blocks_i8(const void* digits, const double* rowscale, int Npad, const double* Kstar,
                                        int ldks, const double* G, const double* W, const double* Bmat,
                                        const double* Ct, double kss, int n, int p, int Q, double* Mk, double* Bk,
                                        void* stream_) {
  return bcbf_posterior_blocks_i8_d(digits, rowscale, Npad, Kstar, ldks, G, W, Bmat, Ct, kss, n, p, Q, Mk, Bk, oz::S,
                                    stream_);
}

extern "C" int bcbf_posterior_blocks_i8_d(const void* digits, const double* rowscale, int Npad, const double* Kstar,
                                          int ldks, const double* G, const double* W, const double* Bmat,
                                          const double* Ct, double kss, int n, int p, int Q, double* Mk, double* Bk,
                                          int ndigits, void* stream_) {
  ::bcbf::ScratchScope scratch_scope(static_cast<cudaStream_t>(stream_));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(ndigits == 6 || ndigits == 7, "bcbf_posterior_blocks_i8: digits=%d (6 or 7)", ndigits);
  BCBF_REQUIRE(digits && rowscale && Kstar && G && Bmat && (Mk || Bk), "bcbf_posterior_blocks_i8: null pointer");
  BCBF_REQUIRE(!Mk || (W && Ct), "bcbf_posterior_blocks_i8: Mk requested without W / Ct");
  BCBF_REQUIRE(n >= 1 && n <= BCBF_MAX_N_DIM, "bcbf_posterior_blocks_i8: n=%d", n);
  BCBF_REQUIRE(Npad > 0 && Npad % oz::TM == 0 && Npad <= oz::kMaxNpad && Q >= 1 && ldks >= Q,
               "bcbf_posterior_blocks_i8: Npad=%d Q=%d ldks=%d", Npad, Q, ldks);
  const int8_t* A = static_cast<const int8_t*>(digits);
#define BCBF_OZ_RUN(P_, D_) oz::run_blocks<P_, D_>(A, rowscale, Npad, Kstar, ldks, G, W, Bmat, Ct, kss, n, Q, Mk, Bk, stream)
  switch (p) {
    case 1: return ndigits == 7 ? BCBF_OZ_RUN(1, 7) : BCBF_OZ_RUN(1, 6);
    case 2: return ndigits == 7 ? BCBF_OZ_RUN(2, 7) : BCBF_OZ_RUN(2, 6);
    case 3: return ndigits == 7 ? BCBF_OZ_RUN(3, 7) : BCBF_OZ_RUN(3, 6);
    case 4: return ndigits == 7 ? BCBF_OZ_RUN(4, 7) : BCBF_OZ_RUN(4, 6);
    default: break;
  }
#undef BCBF_OZ_RUN
  set_last_error("bcbf_posterior_blocks_i8: p=%d unsupported", p);
  return BCBF_ERR_INVALID;
}

extern "C" int bcbf_posterior_var_i8(const void* digits, const double* rowscale, int Npad, const double* Kstar,
                                     int ldks, const double* G, const double* Bmat, double kss, int p, int Q, double* Bk,
                                     void* stream) {
  BCBF_REQUIRE(Bk, "bcbf_posterior_var_i8: null pointer");
  return bcbf_posterior_blocks_i8(digits, rowscale, Npad, Kstar, ldks, G, nullptr, Bmat, nullptr, kss, 1, p, Q, nullptr,
                                  Bk, stream);
}

// Pipeline counters of oz_var_kernel (development aid; adds clock64 reads while enabled): out[0] cycles of the MMA
// issue thread summed over CTAs, [1] of which waiting for operand stages, [2] waiting for the epilogue to drain TMEM,
// [3] producer cycles waiting for a free stage, [4] K steps issued.
// Development aid: oz_var_kernel stops copying the A (bit 0) and / or B (bit 1) digits into shared memory — the MMAs run on
// whatever the stages hold, results are void — to measure what the incoming copies cost the MMA pipeline (tools/oz_sweep.py);
// bit 2: the epilogue drains the accumulators but skips its FP64 work, bit 3: it does not even drain them.
extern "C" int bcbf_oz_debug_skip_loads(int mask) {
  BCBF_REQUIRE(mask >= 0 && mask <= 15, "bcbf_oz_debug_skip_loads: mask %d not in 0..15", mask);
  oz::g_skip = mask;
  return BCBF_OK;
}

extern "C" int bcbf_oz_debug_counters(int enable, unsigned long long out[8]) {
  if (out != nullptr && oz::g_dbg != nullptr) {
    BCBF_CUDA(cudaDeviceSynchronize());
    BCBF_CUDA(cudaMemcpy(out, oz::g_dbg, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  if (enable && oz::g_dbg == nullptr) BCBF_CUDA(cudaMalloc(&oz::g_dbg, 8 * sizeof(unsigned long long)));
  if (oz::g_dbg != nullptr) BCBF_CUDA(cudaMemset(oz::g_dbg, 0, 8 * sizeof(unsigned long long)));
  if (!enable && oz::g_dbg != nullptr) {
    BCBF_CUDA(cudaFree(oz::g_dbg));
    oz::g_dbg = nullptr;
  }
  return BCBF_OK;
}

extern "C" int bcbf_oz_gemm_tn(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
                               double* C, int ldc, int lower, void* stream_) {
  ::bcbf::ScratchScope scratch_scope(static_cast<cudaStream_t>(stream_));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(A && B && C, "bcbf_oz_gemm_tn: null pointer");
  BCBF_REQUIRE(M > 0 && N > 0 && K > 0 && M % oz::TM == 0 && N % oz::TN == 0 && K % oz::KSTEP == 0 && K <= oz::kMaxNpad,
               "bcbf_oz_gemm_tn: M=%d (multiple of 128), N=%d (of 64), K=%d (of 32, <= %d)", M, N, K, oz::kMaxNpad);
  BCBF_REQUIRE(lda >= M && ldb >= N && ldc >= N && ldc % 2 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0,
               "bcbf_oz_gemm_tn: lda=%d ldb=%d ldc=%d (ldc even, C 16-byte aligned)", lda, ldb, ldc);
  BCBF_REQUIRE(!lower || (M == K && N == K), "bcbf_oz_gemm_tn: lower-triangular mode needs square operands");
  return oz::run_gemm_tn(M, N, K, alpha, A, lda, B, ldb, C, ldc, lower ? 1 : 0, stream);
}

extern "C" int bcbf_oz_update(int M, int N, int K, double alpha, const double* PA, int lda, const double* PB, int ldb,
                              double* C, int ldc, int lower, void* stream_) {
  ::bcbf::ScratchScope scratch_scope(static_cast<cudaStream_t>(stream_));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(PA && PB && C, "bcbf_oz_update: null pointer");
  BCBF_REQUIRE(M > 0 && N > 0 && K > 0 && M % oz::TM == 0 && N % oz::TN == 0 && K % oz::KSTEP == 0 && K <= oz::kMaxNpad,
               "bcbf_oz_update: M=%d (multiple of 128), N=%d (of 64), K=%d (of 32, <= %d)", M, N, K, oz::kMaxNpad);
  BCBF_REQUIRE(lda >= K && ldb >= K && ldc >= N && ldc % 2 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0,
               "bcbf_oz_update: lda=%d ldb=%d ldc=%d (ldc even, C 16-byte aligned)", lda, ldb, ldc);
  BCBF_REQUIRE(!lower || N == M, "bcbf_oz_update: lower-triangle mode needs a square C (M=%d, N=%d)", M, N);
  return oz::run_update(M, N, K, alpha, PA, lda, PB, ldb, C, ldc, lower ? 1 : 0, stream);
}

extern "C" int bcbf_oz_update_reserve(int M, int N, int K) {
  void* p = nullptr;
  int rc;
  if ((rc = oz::workspace(9, (size_t)ceil_div(M, oz::TM) * ceil_div(K, oz::KSTEP) * oz::A_STEP, &p))) return rc;
  if ((rc = oz::workspace(10, (size_t)ceil_div(N, oz::TN) * ceil_div(K, oz::KSTEP) * oz::B_STEP, &p))) return rc;
  return oz::workspace(11, sizeof(double) * ((size_t)M + (size_t)N), &p);
}

// Pre-size the operand-digit workspaces of bcbf_oz_gemm (so that a timed caller does not pay cudaMalloc).
extern "C" int bcbf_oz_gemm_reserve(int M, int N, int K) {
  BCBF_REQUIRE(M > 0 && N > 0 && K > 0, "bcbf_oz_gemm_reserve: M=%d N=%d K=%d", M, N, K);
  void* p = nullptr;
  int rc;
  if ((rc = oz::workspace(5, (size_t)ceil_div(M, oz::TM) * ceil_div(K, oz::KSTEP) * oz::A_STEP, &p))) return rc;
  if ((rc = oz::workspace(6, (size_t)ceil_div(N, oz::TN) * ceil_div(K, oz::KSTEP) * oz::B_STEP, &p))) return rc;
  if ((rc = oz::workspace(7, sizeof(double) * (size_t)M, &p))) return rc;
  return oz::workspace(8, 16 * ((size_t)N + (size_t)K), &p);
}

extern "C" int bcbf_oz_gemm(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb,
                            double* C, int ldc, int tri, void* stream_) {
  ::bcbf::ScratchScope scratch_scope(static_cast<cudaStream_t>(stream_));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  BCBF_REQUIRE(A && B && C, "bcbf_oz_gemm: null pointer");
  BCBF_REQUIRE(M > 0 && N > 0 && K > 0 && M % oz::TM == 0 && N % oz::TN == 0 && K % oz::KSTEP == 0 && K <= oz::kMaxNpad,
               "bcbf_oz_gemm: M=%d (multiple of 128), N=%d (of 64), K=%d (of 32, <= %d)", M, N, K, oz::kMaxNpad);
  BCBF_REQUIRE(lda >= K && ldb >= N && ldc >= N && ldc % 2 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0,
               "bcbf_oz_gemm: lda=%d ldb=%d ldc=%d (ldc even, C 16-byte aligned)", lda, ldb, ldc);
  BCBF_REQUIRE(tri >= 0 && tri <= 2 && (tri != oz::kTriGemmALower || M == K) && (tri != oz::kTriGemmBLower || K == N),
               "bcbf_oz_gemm: tri=%d needs a square triangular operand", tri);
  return oz::run_gemm(M, N, K, alpha, A, lda, B, ldb, C, ldc, tri, stream);
}

// CTAs per cluster of oz_var_kernel: 1 (default) or 2 (the pair multicasts the L^-1 digits to each other).
extern "C" int bcbf_oz_set_group(int row_blocks) {
  BCBF_REQUIRE(row_blocks >= 1 && row_blocks <= 32, "bcbf_oz_set_group: 1..32");
  oz::g_group = row_blocks;
  return BCBF_OK;
}

extern "C" int bcbf_oz_set_cluster(int ctas) {
  BCBF_REQUIRE(ctas == 1 || ctas == 2 || ctas == 4, "bcbf_oz_set_cluster: 1, 2 or 4");
  oz::g_cluster = ctas;
  return BCBF_OK;
}

extern "C" int bcbf_oz_profile_enable(int on) {
  for (int i = 0; i < oz::g_prof.n; ++i) {
    cudaEventDestroy(oz::g_prof.e0[i]);
    cudaEventDestroy(oz::g_prof.e1[i]);
  }
  oz::g_prof.n = 0;
  oz::g_prof.on = on != 0;
  return BCBF_OK;
}

extern "C" int bcbf_oz_profile_read(double* total_ms, int* launches) {
  double tot = 0.0;
  for (int i = 0; i < oz::g_prof.n; ++i) {
    BCBF_CUDA(cudaEventSynchronize(oz::g_prof.e1[i]));
    float ms = 0.f;
    BCBF_CUDA(cudaEventElapsedTime(&ms, oz::g_prof.e0[i], oz::g_prof.e1[i]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = oz::g_prof.n;
  return BCBF_OK;
}
