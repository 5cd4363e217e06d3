/*
 * bcbf.h — C ABI of the B200-native MVGP hot path (libbcbf.so).
 *
 * Drop-in boundary for the matrix-variate GP path of wecacuee/Bayesian_CBF.  The reference has no
 * FFI: the path sits behind Python classes (bayes_cbf/control_affine_model.py).  Each entry point below
 * names the reference statement(s) it replaces (file:line in the reference repo).  The Python host side
 * (bayesian_cbf_b200/) keeps the reference's class API and binds these symbols with ctypes; see
 * INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - all matrices are row-major float64; "device" pointers are CUDA device pointers on the current device;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous w.r.t. the
 *     host unless stated otherwise;
 *   - N  = number of training points, Npad = N rounded up to a multiple of 128 (bcbf_padded()); factor-sized
 *     buffers are Npad x Npad with leading dimension ld >= Npad, the pad region holds the identity;
 *   - n  = state dim, m = control dim, p = 1 + m (homogeneous control [1;u]);
 *   - the library keeps per-device scratch buffers (posterior partial sums, operand digits of the int8 kernels) and, for
 *     bcbf_potrf, a set of look-ahead streams.  Entry points that use them (posterior_*, oz_*, potrf, trtri, model_*) may
 *     be called on different streams and from different host threads of one device: a per-device lock is held while a call
 *     enqueues its work, and a call on another stream than its predecessor first waits, on the GPU, for that predecessor
 *     (calls that share scratch execute in call order; they do not overlap each other).  Calls on ONE stream are ordered by
 *     the stream as usual.  A bcbf_model handle is used by one thread at a time;
 *   - return value: BCBF_OK (0) or a negative error code; bcbf_last_error() describes the last failure
 *     of the calling thread.  There is NO CPU fallback anywhere: without a CUDA device every compute
 *     entry point returns BCBF_ERR_CUDA.
 */
#ifndef BCBF_H_
#define BCBF_H_

#ifdef __cplusplus
extern "C" {
#endif

#define BCBF_OK 0
#define BCBF_ERR_INVALID (-1)
#define BCBF_ERR_CUDA (-2)
#define BCBF_ERR_NOT_PD (-3) /* Cholesky met a non-positive pivot: the caller retries with more jitter */
#define BCBF_ERR_NOT_FITTED (-4)

#define BCBF_BLOCK 128
#define BCBF_MAX_N_DIM 8 /* state dim n  */
#define BCBF_MAX_P_DIM 4 /* p = 1 + m    */

const char* bcbf_last_error(void);
int bcbf_version(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches claim). */
unsigned long long bcbf_launch_count(void);
/* CUDA-event timing of the dominant kernel (post_var_kernel, the N^2 p contraction): enable clears the history;
 * read synchronises and returns the summed device time and the number of launches since enable.            */
int bcbf_profile_enable(int on);
int bcbf_profile_read(double* total_ms, int* launches);
/* Development aid: pipeline counters of post_var_kernel (producer / consumer barrier wait cycles); see posterior.cu. */
int bcbf_debug_counters(int enable, unsigned long long out[8]);
/* N rounded up to the block size the kernels tile by. */
int bcbf_padded(int N);
/* Bytes of scratch bcbf_potrf/bcbf_trtri need for a factor of padded size Npad. */
long long bcbf_dinv_elems(int Npad);

/* ------------------------------------------------------------------------------------------------
 * (1) Control-affine Gram matrix.
 * Replaces control_affine_model.py:370-372:  KXX = k(X,X); uBu = UH @ B @ UH.T; Kb = KXX * uBu
 * with k = ScaleKernel(RBFKernel(ard))  (control_affine_model.py:164-171).
 *   X (N,n)  UH (N,p)  Bmat (p,p)  lengthscale (n)  -> Kb (Npad,Npad; ld)
 * Writes the FULL symmetric matrix on [0,N)^2 (so the result equals the reference's dense Kb), the identity on
 * the pad diagonal and zeros elsewhere in the pad.
 */
int bcbf_gram_train(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                    double outputscale, int N, int n, int p, double* Kb, int ld, int Npad, void* stream);

/* The same, producing only what the factorisation reads: the 64 x 64 tiles on or below the diagonal (diagonal tiles whole)
 * and the identity on the pad diagonal; storage above those tiles is left untouched.  Half the exps and half the bytes
 * of bcbf_gram_train; entry (i, j), j <= i, has the same bits in both.  Npad must be a multiple of 64.  (Fit path.)   */
int bcbf_gram_train_lower(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                          double outputscale, int N, int n, int p, double* Kb, int ld, int Npad, void* stream);

/* Residual of the alpha solve against the matrix that was factorised, in compensated arithmetic:
 *     R[:, :nc] = Y - (Kb + jitter_scale * diag(jitter)) alpha
 * Kb is NOT read from memory (bcbf_potrf overwrote it with L): every entry is re-evaluated from X, UH, B exactly as
 * bcbf_gram_train[_lower] produced it (bit-identical, diagonal jitter added with the same single rounding as bcbf_potrf),
 * and the products are accumulated error-free (TwoProduct/TwoSum, Dot2), i.e. as if in ~106-bit arithmetic rounded once.
 * alpha (N, lda), Y (N, ldy), R (N, ldr), nc <= 8 columns; scratch >= bcbf_gram_resid_scratch_elems(N) doubles.
 * This is what lets the iterative refinement of bcbf_alpha_refine converge to the FP64 rounding of the exact solution
 * although cond(Kb) ~ 1e10..1e13 (reference: alpha = torch.cholesky_solve(Y, L), control_affine_model.py:545).       */
long long bcbf_gram_resid_scratch_elems(int N);
int bcbf_gram_resid(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                    double outputscale, int N, int n, int p, const double* jitter, double jitter_scale,
                    const double* alpha, int lda, const double* Y, int ldy, int nc, double* R, int ldr, double* scratch,
                    long long scratch_elems, void* stream);
/* The same residual with Kb READ from memory: Kb (N,N; ldk) with its lower triangle as bcbf_gram_train_lower wrote it (the
 * upper triangle is not touched: entry (i,j), j > i, is read as (j,i)).  Same accumulation order, so the same result
 * bits as bcbf_gram_resid, at the cost of streaming 2 N^2 doubles instead of evaluating N^2 exponentials.             */
int bcbf_gram_resid_stored(const double* Kb, int ldk, int N, const double* jitter, double jitter_scale,
                           const double* alpha, int lda, const double* Y, int ldy, int nc, double* R, int ldr,
                           double* scratch, long long scratch_elems, void* stream);

/* Cross Gram  Kstar[i, j] = k(X_i, Xq_j)   (control_affine_model.py:536 / :1051, the k_xs / k_sx factor).
 *   X (N,n), Xq (Q,n) -> Kstar (Npad, ldks) row-major with ldks >= Q; rows >= N are written as zero.     */
int bcbf_cross_gram(const double* X, const double* Xq, const double* lengthscale, double outputscale, int N,
                    int Q, int n, double* Kstar, int ldks, int Npad, void* stream);

/* General control-affine Gram between two point sets (no padding):
 *   out[i, j] = k(x1_i, x2_j) * (uh1_i^T B uh2_j)        (a, c; ld >= c)
 * Replaces kb* = k_xs(Xtrain, Xtest) * (UHtrain @ B @ UHtest.t()) and kb** = k_ss(Xtest, Xtestp) * (UHtest @ B @
 * UHtestp.t()) of custom_predict (control_affine_model.py:536, :549-553) and, with one-hot uh2 rows, the
 * frakB(x) matrix of the Exact class (:1051).  UH1 = UH2 = NULL gives the plain data kernel k(X1, X2).    */
int bcbf_gram_ca(const double* X1, const double* UH1, int a, const double* X2, const double* UH2, int c,
                 const double* Bmat, const double* lengthscale, double outputscale, int n, int p, double* out,
                 int ld, void* stream);

/* Control-affine weighting of a data-kernel matrix evaluated elsewhere:  out[i,j] = K[i,j] * (uh1_i^T B uh2_j).
 * K (a,c; ldk), UH1 (a,p), UH2 (c,p), out (a,c; ldo; may alias K).  The plug-in path of HetergeneousMatrixVariateKernel
 * (matrix_variate_multitask_kernel.py:196-204: any data_covar_module; kernel1 / correlation_kernel_12 :112-134) — the
 * stock RBF-ARD x scale kernel takes the fused bcbf_gram_ca instead.                                                    */
int bcbf_ca_weight(const double* K, int ldk, const double* UH1, int a, const double* UH2, int c, const double* Bmat,
                   int p, double* out, int ldo, void* stream);

/* General k(X1, X2) (a x c) dense, plus optional closed-form derivative blocks
 *   dK[i,j,:]   = d k(x1_i, x2_j) / d x1_i                      (a,c,n)     [may be NULL]
 *   d2K[i,j,:,:] = d^2 k(x1_i, x2_j) / d x1_i d x2_j^T           (a,c,n,n)   [may be NULL]
 * Replaces the autograd-through-gpytorch derivative kernels grad_ksx / grad_kxs / Hessian_kxx
 * (control_affine_model.py:465-477, misc.py:236-245).                                                   */
int bcbf_rbf_blocks(const double* X1, const double* X2, const double* lengthscale, double outputscale, int a,
                    int c, int n, double* K, double* dK, double* d2K, void* stream);

/* Hyper-parameter gradient of the train Gram matrix (backward of bcbf_gram_train; the fit path).
 * With P = Kb^-1 (N,N; ldp), alpha = P Y (N,nout; lda) and alphaAi = alpha A^-1, the adjoint of the MVGP log marginal
 * likelihood w.r.t. Kb is Gbar = 1/2 (alphaAi alpha^T - nout P)  (SURVEY 8a-13); this accumulates
 *     out[0] = sum Gbar dKb/d outputscale,  out[1+d] = sum Gbar dKb/d lengthscale_d  (d < n),
 *     out[1+BCBF_MAX_N_DIM + a*BCBF_MAX_P_DIM + b] = sum Gbar dKb/dB_ab
 * deterministically (per-tile partials in `partial`, >= ceil(N/64)^2 * out_elems doubles, then a fixed-order sum).
 * Replaces autograd through ExactMarginalLogLikelihood (control_affine_model.py:309-325).                        */
int bcbf_gram_train_backward(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                             double outputscale, int N, int n, int p, const double* Pinv, int ldp,
                             const double* alphaAi, const double* alpha, int lda, int nout, double* partial,
                             long long partial_elems, double* out, void* stream);
int bcbf_gram_backward_layout(int* out_elems, int* max_n, int* max_p);

/* ------------------------------------------------------------------------------------------------
 * (2) Blocked FP64 Cholesky  A + scale*diag(jitter) = L L^T, in place, lower (upper triangle is zeroed).
 * Replaces make_psd's  torch.linalg.cholesky(Kb + factor * eye * rand)  (control_affine_model.py:907-911).
 *   A (Npad,Npad; ld) in/out;  jitter (N) may be NULL;  dinv: bcbf_dinv_elems(Npad) doubles of scratch that
 *   receives the inverses of the 128x128 diagonal blocks of L;  info: device int, 0 on success else
 *   1 + index of the first non-positive pivot (LAPACK convention).  Asynchronous: read *info after a sync,
 *   or call bcbf_check_info() which synchronises the stream and maps info != 0 to BCBF_ERR_NOT_PD.
 */
int bcbf_potrf(double* A, int ld, int Npad, int N, const double* jitter, double jitter_scale, double* dinv,
               int* info, void* stream);
int bcbf_check_info(const int* info, void* stream);
/* The large trailing updates of bcbf_potrf (single matrix, >= 1024 rows left) run on the int8 tensor cores
 * (bcbf_oz_update) by default; 0 keeps them on the FP64 pipe. */
int bcbf_set_potrf_i8(int on);
/* Diagonal-block kernel of bcbf_potrf[_batched]: 1 = potf2_inv2_kernel (default: diagonal factorisations overlapped with
 * the trailing update inside the CTA, inverse by recursive doubling), 0 = round 1's potf2_inv_kernel (kept for A/B runs). */
int bcbf_set_potf2_variant(int v);

/* Linv = L^{-1} (lower; strictly-upper blocks are zero) from L and the diagonal-block inverses of bcbf_potrf.
 * scratch: Npad*Npad doubles.  Everything downstream (alpha, v = L \ kb*, posterior covariance) multiplies by
 * Linv instead of running triangular solves (torch.cholesky_solve / torch.linalg.solve at
 * control_affine_model.py:545,565,575,1053).                                                              */
int bcbf_trtri(const double* L, const double* dinv, double* Linv, double* scratch, int ld, int Npad,
               void* stream);
/* The levels of the divide-and-conquer inverse with half-size >= 2048 run on the int8 tensor cores (bcbf_oz_gemm) by
 * default; 0 keeps every level on the FP64 pipe. */
int bcbf_set_trtri_i8(int on);

/* C (M,Ncols; ldc) = alpha * op(A) * B + beta * C  with A = a lower-triangular factor-sized matrix
 * (Npad,Npad; lda): op = identity (trans=0) or transpose (trans=1).  B (Npad,Ncols; ldb).
 * Used for  v = Linv @ kb*,  alpha = Linv^T (Linv Y).                                                  */
int bcbf_trmm_lower(const double* A, int lda, int Npad, int trans, const double* B, int ldb, int ncols,
                    double alpha, double beta, double* C, int ldc, void* stream);

/* alpha = (Kb + jitter_scale diag(jitter))^-1 Y  (control_affine_model.py:545, alpha = cholesky_solve(Y, L)):
 * alpha0 = Linv^T (Linv Y) followed by `iters` steps of iterative refinement  alpha += Linv^T Linv (Y - Kb' alpha)  with
 * the residual of bcbf_gram_resid.  Linv (Npad,Npad; ld) from bcbf_trtri; Y and alpha (Npad, ldy) with zero pad rows,
 * nc <= ldy <= 8 columns in use (ldy even); jitter (N) / jitter_scale as given to bcbf_potrf (jitter may be NULL);
 * scratch >= bcbf_alpha_refine_scratch_elems(N, Npad, ldy) doubles.  Each step gains 1.5-2 digits in the posterior mean
 * at the bench shapes (cond(Kb) ~ 1e11); iters = 3 (what bcbf_model_fit and the host class run) is converged at
 * N = 16384 (a fourth step moves the mean by < 1e-12); iters = 0 is the plain explicit-inverse product.              */
long long bcbf_alpha_refine_scratch_elems(int N, int Npad, int ldy);
int bcbf_alpha_refine(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                      double outputscale, int N, int n, int p, const double* jitter, double jitter_scale,
                      const double* Linv, int ld, int Npad, const double* Y, int ldy, int nc, int iters, double* alpha,
                      double* scratch, long long scratch_elems, void* stream);
/* bcbf_alpha_refine with a workspace kb_ws (Npad,Npad; ldk >= Npad) or NULL: Kb is written there once (bcbf_gram_train_lower)
 * and the `iters` residuals read it (bcbf_gram_resid_stored) instead of re-evaluating it — same alpha, bit for bit.    */
int bcbf_alpha_refine_ws(const double* X, const double* UH, const double* Bmat, const double* lengthscale,
                         double outputscale, int N, int n, int p, const double* jitter, double jitter_scale,
                         const double* Linv, int ld, int Npad, const double* Y, int ldy, int nc, int iters, double* alpha,
                         double* scratch, long long scratch_elems, double* kb_ws, int ldk, void* stream);

/* Row-major C(M,N) = alpha * op(A) op(B) + beta * C on the FP64 tensor-core GEMM.  op(A) is M x K: transa = 0 ->
 * A stored (M,K), 1 -> stored (K,M);  op(B) is K x N: transb = 0 -> stored (K,N), 1 -> stored (N,K).  M, N, K and
 * the leading dimensions must be even and the pointers 16-byte aligned (callers zero-pad).  Replaces the dense
 * products of custom_predict: kb*^T alpha (:547), v^T v' (:586), kb*^T.reshape(bp,N) @ Bdagger (:1079-1088).   */
int bcbf_gemm(int transa, int transb, int M, int N, int K, double alpha, const double* A, int lda, const double* B,
              int ldb, double beta, double* C, int ldc, void* stream);
/* Tile shape of the FP64 GEMM: 0 = by problem size (default: 32 x 128 row tiles when the problem has at most 74 tiles of
 * 128 x 128, i.e. would leave most SMs idle), 1 = always 128 x 128, 2 = always 32 x 128.  Results are identical to the
 * last bit either way (same k order per output element); the knob exists for tests and timing.                        */
int bcbf_set_gemm_tile_policy(int policy);

/* R independent products, element strides sA / sB / sC between consecutive operands (even). */
int bcbf_gemm_batched(int transa, int transb, int M, int N, int K, double alpha, const double* A, int lda, long long sA,
                      const double* B, int ldb, long long sB, double beta, double* C, int ldc, long long sC, int R,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3) Batched posterior of F(x) over many query states (matrix form).
 * Replaces ControlAffineRegressorExact._custom_predict_matrix, per-query diagonal blocks
 * (control_affine_model.py:1051-1091):
 *     M_k(x) = C^T + Y^T Kb^{-1} frakB(x)                       (Q,n,p)
 *     B_k(x) = B k(x,x) - frakB(x)^T Kb^{-1} frakB(x)           (Q,p,p)
 * with frakB(x) = k(X,x)[:,None] * G,  G = UH @ B (Npad,p; pad rows zero),  W[i, r*p+q] = alpha[i,r]*G[i,q]
 * (Npad, n*p).  Kstar from bcbf_cross_gram.  Ct is C^T (n,p).  kss = k(x,x) = outputscale.
 * The random output jitter the reference adds at :1089 is NOT applied here (host adds it when asked).
 */
int bcbf_posterior_blocks(const double* Linv, int ld, int Npad, const double* Kstar, int ldks, const double* G,
                          const double* W, const double* Bmat, const double* Ct, double kss, int n, int p, int Q,
                          double* Mk, double* Bk, void* stream);

/* u-contraction epilogue (control_affine_model.py:952-958):  mean = M_k [1;u]  (Q,n),
 * svar = [1;u]^T B_k [1;u] (Q).  UHq (Q,p).                                                             */
int bcbf_contract_u(const double* Mk, const double* Bk, const double* UHq, int n, int p, int Q, double* mean,
                    double* svar, void* stream);

/* Fold-in form of the base class (control_affine_model.py:536-586): one column per query,
 *     mean = UHq C + kb*^T alpha (Q,n),   svar = kb** - |Linv kb*|^2 (Q)
 * with kb*[i] = k(X_i,x) * (G_i . uh).  Uses Kstar from bcbf_cross_gram.  alpha (Npad,n; pad rows zero).  */
int bcbf_posterior_fu(const double* Linv, int ld, int Npad, const double* Kstar, int ldks, const double* G,
                      const double* alpha, const double* Bmat, const double* C, const double* UHq, double kss,
                      int n, int p, int Q, double* mean, double* svar, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (3b) The same B_k(x) with the N^2 p contraction on the int8 tensor cores (tcgen05), FP64-accurate: both operands are
 * split error-free into seven signed 8-bit digits (Ozaki-type splitting, csrc/ozaki.cu), all integer arithmetic is exact,
 * the FP64 value is recombined in the kernel epilogue.  Same inputs and outputs as the B_k half of
 * bcbf_posterior_blocks (control_affine_model.py:1051-1091); results agree with the DMMA path to FP64 rounding level.
 *   bcbf_oz_factor_bytes(Npad): size of the digit array of L^-1;  bcbf_oz_max_npad(): largest supported Npad (exact int32
 *   accumulation);  bcbf_oz_split_factor: Linv (Npad,Npad; ld) -> digits (bcbf_oz_factor_bytes bytes), rowscale (Npad);
 *   once per fit.  bcbf_posterior_var_i8: Kstar from bcbf_cross_gram (ldks >= Q), G (Npad,p) -> Bk (Q,p,p).           */
long long bcbf_oz_factor_bytes(int Npad);
int bcbf_oz_max_npad(void);
int bcbf_oz_split_factor(const double* Linv, int ld, int Npad, void* digits, double* rowscale, void* stream);
int bcbf_posterior_var_i8(const void* digits, const double* rowscale, int Npad, const double* Kstar, int ldks,
                          const double* G, const double* Bmat, double kss, int p, int Q, double* Bk, void* stream);
/* M_k and B_k together (either may be NULL): the posterior mean partial sums ride on the pass over K* that finds the
 * column scales of frakB.  Arguments as bcbf_posterior_blocks.                                                    */
int bcbf_posterior_blocks_i8(const void* digits, const double* rowscale, int Npad, const double* Kstar, int ldks,
                             const double* G, const double* W, const double* Bmat, const double* Ct, double kss, int n,
                             int p, int Q, double* Mk, double* Bk, void* stream);
/* The same with an explicit digit count (6 or 7; the entry points above use 7).  Seven digits keep the 28 digit products
 * whose weights reach 2^-56 of row scale x column scale (FP64 rounding level: B_k to ~4e-13 of the prior scale at
 * N = 16384).  Six digits keep 21 products (2^-48: B_k to ~3e-11 of the prior scale, still inside the 1e-9 parity
 * tolerance) for 25 % less tensor work — an opt-in trade, never the default.  The digit array of L^-1 must have been split
 * with the same count (bcbf_oz_factor_bytes_d bytes).                                                                   */
long long bcbf_oz_factor_bytes_d(int Npad, int digits);
int bcbf_oz_split_factor_d(const double* Linv, int ld, int Npad, void* digits_out, double* rowscale, int digits,
                           void* stream);
int bcbf_posterior_blocks_i8_d(const void* digits_in, const double* rowscale, int Npad, const double* Kstar, int ldks,
                               const double* G, const double* W, const double* Bmat, const double* Ct, double kss, int n,
                               int p, int Q, double* Mk, double* Bk, int digits, void* stream);
/* oz_var_kernel runs as single CTAs (1, default) or as clusters of 2 CTAs that multicast the L^-1 digits to each other
 * (bit-identical results; measured no faster: the shared-memory port, not L2, is the limiter). */
int bcbf_oz_set_cluster(int ctas);
/* Work order of oz_var_kernel: consecutive work items sweep `row_blocks` row blocks of L^-1 x consecutive column tiles, so
 * the 148 CTAs running side by side share row_blocks digit blobs of L^-1 and 148 / row_blocks of frakB through L2
 * (28 row_blocks + 14 * 148 / row_blocks KB per K step: minimal near 8).  Results do not depend on it (bit-identical). */
int bcbf_oz_set_group(int row_blocks);
/* Development aid: pipeline counters of oz_var_kernel (see csrc/ozaki.cu). */
int bcbf_oz_debug_counters(int enable, unsigned long long out[8]);
/* Development aid (timing experiments only; results are void while set): oz_var_kernel skips the shared-memory copies of
 * the L^-1 digits (bit 0) and / or the frakB digits (bit 1).  0 restores normal operation.                              */
int bcbf_oz_debug_skip_loads(int mask);
/* General FP64-accurate GEMM on the int8 tensor cores (same digit splitting; csrc/ozaki.cu: oz_gemm_kernel):
 *   C (M,N; ldc) = alpha * A (M,K; lda) * B (K,N; ldb),  row-major, M % 128 == 0, N % 64 == 0, K % 32 == 0, K <= bcbf_oz_max_npad();
 *   tri = 0, or 1: A is square lower triangular (its strictly upper storage is not read), or 2: B is square lower triangular.
 * bcbf_trtri uses it for the large levels of the triangular inverse.                                                  */
int bcbf_oz_gemm(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C, int ldc,
                 int tri, void* stream);
/* C (M,N; ldc) = alpha * A^T B with A (K,M; lda), B (K,N; ldb) row-major, same size rules; lower != 0: A and B are square
 * lower triangular (Kb^-1 = L^-T L^-1 of the log-marginal gradient; replaces the dense product in mll.py for N >= 2048). */
int bcbf_oz_gemm_tn(int M, int N, int K, double alpha, const double* A, int lda, const double* B, int ldb, double* C,
                    int ldc, int lower, void* stream);
/* Rank-K update on the int8 tensor cores (oz_update_kernel):  C (M,N; ldc) += alpha * PA (M,K; lda) * PB (N,K; ldb)^T,
 * row-major, M % 128 == 0, N % 64 == 0, K % 32 == 0; lower != 0 (M == N): only the 128 x 64 tiles that touch the lower
 * triangle are updated (whole tiles, so the part of a diagonal tile above the diagonal receives the symmetric values).
 * bcbf_potrf uses it for the trailing update A22 -= L21 L21^T.                                                       */
int bcbf_oz_update(int M, int N, int K, double alpha, const double* PA, int lda, const double* PB, int ldb, double* C,
                   int ldc, int lower, void* stream);
int bcbf_oz_update_reserve(int M, int N, int K);
/* Pre-size bcbf_oz_gemm's internal workspaces for products up to (M,K) x (K,N). */
int bcbf_oz_gemm_reserve(int M, int N, int K);
/* CUDA-event timing of oz_var_kernel launches (bench.py's roofline leg), like bcbf_profile_enable/read. */
int bcbf_oz_profile_enable(int on);
int bcbf_oz_profile_read(double* total_ms, int* launches);

/* Relative-degree-1 control-barrier-condition terms in closed form, replacing the autograd
 * extraction of cbc2_quadratic_terms (cbc2.py:7-23) + convert_cbc_terms_to_socp_terms
 * (unicycle_move_to_pose.py:837-878), batched over Q constraints:
 *   row = grad_h^T (Fbar + M_k);  e = row[0] + gamma*h;  bfe = row[1:]
 *   Asq = (grad_h^T A grad_h) * B_k = Ls Ls^T;  A_socp = Ls^T[:,1:] (p,m);  bfb = Ls^T[:,0] (p)
 * Fbar (Q,n,p) may be NULL.  Outputs: bfe (Q,m), e (Q), Asq (Q,p,p), A_socp (Q,p,m), bfb (Q,p),
 * status (Q) = 0 or 1+index of a non-positive pivot of Asq.                                             */
int bcbf_cbc1_terms(const double* Mk, const double* Bk, const double* Amat, const double* grad_h, const double* h,
                    const double* Fbar, double gamma, int n, int p, int Q, double* bfe, double* e, double* Asq,
                    double* A_socp, double* bfb, int* status, void* stream);

/* Batched  Asq (Q,p,p) = Ls Ls^T,  A_socp = Ls^T[:,1:] (Q,p,m),  bfb = Ls^T[:,0] (Q,p): the Cholesky inside
 * convert_cbc_terms_to_socp_terms (controllers.py:446-451, unicycle_move_to_pose.py:861).  reg > 0 enables the
 * reference's singular fallback (retry once with Asq + reg I; controllers.py:447-449 uses 1e-3).
 * status (Q, may be NULL) = 0 or 1 + index of the non-positive pivot of the last attempt.                   */
int bcbf_socp_factor(const double* Asq, int p, int Q, double reg, double* A_socp, double* bfb, int* status,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * (4) Ensembles of R small independent MVGPs — one per rollout, each with its own training set (N points, equal N),
 * hyper-parameters and factor (BASELINE configs[4]).  All arrays are device pointers, rollout-major:
 *   X (R,N,n)  UH (R,N,p)  Xdot (R,N,n)  lengthscale (R,n)  outputscale (R)  Bmat (R,p,p)  C (R,p,n)
 *   factor-sized arrays (R,Npad,Npad) with ld = Npad;  dinv (R, bcbf_dinv_elems(Npad));  info int[R].
 * Batched twins of (1)-(2): every rollout's Gram / Cholesky / inverse / alpha in the same launches.
 * Per control step each rollout evaluates custom_predict at ONE state (unicycle_move_to_pose.py:880-920 ->
 * control_affine_model.py:931-961, 983-1096); bcbf_ens_posterior does that for all rollouts in one launch, streaming
 * each rollout's own L^-1 (lower triangle, 4 N^2 bytes) from HBM: the bandwidth-bound regime of SURVEY 8d.       */
int bcbf_ens_gram(const double* X, const double* UH, const double* lengthscale, const double* outputscale,
                  const double* Bmat, int R, int N, int n, int p, double* Kb, int Npad, void* stream);
int bcbf_potrf_batched(double* A, int ld, int Npad, int N, const double* jitter /* (R,N) or NULL */,
                       double jitter_scale, double* dinv, int* info, int R, void* stream);
int bcbf_trtri_batched(const double* L, const double* dinv, double* Linv, double* scratch, int ld, int Npad, int R,
                       void* stream);
int bcbf_trmm_lower_batched(const double* A, int lda, int Npad, int trans, const double* B, int ldb, int ncols,
                            double alpha, double beta, double* C, int ldc, int R, void* stream);
/* G = UH B (R,Npad,p; pad rows 0),  Y = Xdot - UH C (R,Npad,ldy; pad 0)            (control_affine_model.py:525-532) */
int bcbf_ens_prep(const double* UH, const double* Xdot, const double* Bmat, const double* C, int R, int N, int Npad,
                  int n, int p, int ldy, double* G, double* Y, void* stream);
/* W[r,i,c*p+j] = alpha[r,i,c] * G[r,i,j]   (R,Npad,n*p) */
int bcbf_ens_w(const double* alpha, int ldy, const double* G, int R, int Npad, int n, int p, double* W, void* stream);
/* Batched twin of bcbf_gram_train_backward: out (R, 1 + BCBF_MAX_N_DIM + BCBF_MAX_P_DIM^2) per rollout
 * [d/d outputscale | d/d lengthscale | d/dB]; Pinv (R,Npad,Npad), alphaAi / alpha (R,N,nout) contiguous;
 * partial: >= R * ceil(N/64)^2 * (1 + BCBF_MAX_N_DIM + BCBF_MAX_P_DIM^2) doubles.  (Per-rollout hyper-parameter
 * refits of the learning rollouts, unicycle_move_to_pose.py:359-386.)                                              */
int bcbf_ens_gram_backward(const double* X, const double* UH, const double* lengthscale, const double* outputscale,
                           const double* Bmat, const double* Pinv, const double* alphaAi, const double* alpha, int R,
                           int N, int Npad, int n, int p, int nout, double* partial, long long partial_elems,
                           double* out, void* stream);
/* At[r] = A[r]^T for R square (Npad,Npad) matrices (Npad multiple of 32; out of place).                          */
int bcbf_ens_transpose(const double* A, double* At, int Npad, int R, void* stream);
/* xq (R,n): one query state per rollout -> Mk (R,n,p), Bk (R,p,p) (no output jitter).  LinvT is the TRANSPOSED
 * inverse factor (bcbf_ens_transpose of bcbf_trtri_batched's output): row k holds L^-1[:, k].                     */
int bcbf_ens_posterior(const double* LinvT, const double* X, const double* G, const double* W,
                       const double* lengthscale, const double* outputscale, const double* Bmat, const double* C,
                       const double* xq, int R, int N, int Npad, int n, int p, double* Mk, double* Bk, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (5) Batched tiny second-order-cone programs — the safety program of one control step, for Q rollouts at once:
 *     minimise sum_i w_i (y_i - r_i)^2   s.t.   c_k^T y + d_k >= rho ||A_k y + b_k||_2,  k < K
 * y (Q,nv) with nv <= 4 ([relaxation; u]), K <= 4 cones of dimension pc <= 4.  w is (nv) shared (w_per_problem = 0) or
 * (Q,nv); r (Q,nv) or NULL (zeros); c (Q,K,nv), d (Q,K), A (Q,K,pc,nv), b (Q,K,pc).  status (Q): 0 optimal, 1 infeasible
 * (then y = NaN) — the decision on which the reference raises ValueError(problem.status).  Replaces the cvxpy + GUROBI
 * solve of ControllerCLFBayesian.control (unicycle_move_to_pose.py:926-964) / optimizers.py:91-116.
 * Log-barrier interior point, float64, deterministic; tol = duality-gap tolerance on the objective (e.g. 1e-9).   */
int bcbf_socp_solve(int Q, int nv, int K, int pc, double rho, const double* w, int w_per_problem, const double* r,
                    const double* c, const double* d, const double* A, const double* b, double tol, double* y,
                    int* status, int* iters /* may be NULL */, void* stream);
/* The same with a linear term:  minimise sum_i w_i (y_i - r_i)^2 + q^T y  (q (Q,nv) or NULL; w may be all zero: a pure
 * linear objective, the call shape of optimizers.py:6-116 — optimizer_socp_cvxopt / optimizer_socp_cvxpy(u0,
 * linear_objective, [(name, (A, bfb, bfc, d)), ...]); cones of fewer than pc rows are padded with zero rows).  Pinned by
 * the reference's own known answer, tests/test_optimizers.py:6-119 (the cvxopt documentation SOCP).                  */
int bcbf_socp_solve_lin(int Q, int nv, int K, int pc, double rho, const double* w, int w_per_problem, const double* r,
                        const double* q, const double* c, const double* d, const double* A, const double* b, double tol,
                        double* y, int* status, int* iters /* may be NULL */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Model handle: owns device memory for one fitted MVGP; HOST-pointer interface (pinned or pageable).
 * This is what a non-torch caller (and bench.py's e2e leg) binds.
 */
typedef struct bcbf_model bcbf_model;

typedef struct bcbf_hyper {
  int n;                                              /* state dim */
  int p;                                              /* 1 + control dim */
  double outputscale;                                 /* ScaleKernel.outputscale */
  double lengthscale[BCBF_MAX_N_DIM];                 /* RBFKernel.lengthscale (ARD) */
  double A[BCBF_MAX_N_DIM * BCBF_MAX_N_DIM];          /* task_covar.U.covar_matrix (n,n) row-major */
  double B[BCBF_MAX_P_DIM * BCBF_MAX_P_DIM];          /* task_covar.V.covar_matrix (p,p) row-major */
  double C[BCBF_MAX_P_DIM * BCBF_MAX_N_DIM];          /* mean constants (p,n) row-major */
} bcbf_hyper;

int bcbf_model_create(bcbf_model** out, int device);
void bcbf_model_destroy(bcbf_model* m);
/* Gram + jittered Cholesky + Linv + alpha at fixed hyper-parameters ("fit time" of BASELINE.json; what
 * _perturbed_cholesky + the alpha solve of custom_predict do on first use, control_affine_model.py:366-385,545).
 * X (N,n) U (N,m) Xdot (N,n) jitter (N) are HOST pointers.  Synchronous.  BCBF_ERR_NOT_PD -> retry with 10x scale. */
int bcbf_model_fit(bcbf_model* m, const bcbf_hyper* hyp, const double* X, const double* U, const double* Xdot,
                   int N, const double* jitter, double jitter_scale);
/* Posterior of F(x)[1;u] for Q queries, HOST pointers in and out (any output may be NULL):
 *   mean (Q,n), svar (Q) [cov = svar * A], Mk (Q,n,p), Bk (Q,p,p).  Synchronous; copies are inside the call. */
int bcbf_model_query(bcbf_model* m, const double* Xq, const double* Uq, int Q, double* mean, double* svar,
                     double* Mk, double* Bk);
/* Same with DEVICE pointers on `stream`, asynchronous (used by the sharded bench and the Python host). */
int bcbf_model_query_device(bcbf_model* m, const double* Xq, const double* Uq, int Q, double* mean, double* svar,
                            double* Mk, double* Bk, void* stream);
/* Multi-GPU: one rank fits, the others bcbf_model_alloc_state, every rank reads the device pointers of the state with
 * bcbf_model_state (a pure getter: it changes nothing), the buffers are broadcast (NCCL), and the receiving ranks call
 * bcbf_model_adopt, which marks the handle fitted (BCBF_ERR_INVALID without allocated state).  Layout: DESIGN.md.      */
int bcbf_model_state(bcbf_model* m, int* N, int* Npad, double** L, double** Linv, double** alpha, double** G,
                     double** W, double** Xtrain);
int bcbf_model_alloc_state(bcbf_model* m, const bcbf_hyper* hyp, int N);
int bcbf_model_adopt(bcbf_model* m);
/* The lower-triangular 128-blocks of a factor-sized matrix as one contiguous vector (block row after block row, row i =
 * a (128, 128 (i+1)) row-major matrix): what is broadcast of L^-1, 4 Npad (Npad + 128) bytes instead of 8 Npad^2.
 * bcbf_unpack_lower also zeroes the strictly-upper blocks of the destination.                                      */
long long bcbf_packed_lower_elems(int Npad);
int bcbf_pack_lower(const double* M, int ld, int Npad, double* buf, void* stream);
int bcbf_unpack_lower(const double* buf, int Npad, double* M, int ld, void* stream);
/* Milliseconds spent in the stages of the last bcbf_model_fit (gram, potrf, trtri, alpha, total). */
int bcbf_model_fit_timing(bcbf_model* m, double out_ms[5]);
/* Which kernel computes B_k in bcbf_model_query*: 0 = FP64 tensor pipe (DMMA, post_var_kernel), 1 = int8 tensor cores
 * with error-free digit splitting (oz_var_kernel; larger factors than bcbf_oz_max_npad() silently stay on path 0: same
 * results).  With path 1 the fit also splits
 * L^-1 into digits (bcbf_model_oz_split_ms: device time of that step in the last fit; it is part of fit "total").  */
int bcbf_model_set_var_path(bcbf_model* m, int path);
/* Digits per operand of path 1: 7 (default) or 6 (see bcbf_posterior_blocks_i8_d); takes effect at the next query. */
int bcbf_model_set_oz_digits(bcbf_model* m, int digits);
int bcbf_model_get_var_path(bcbf_model* m);
double bcbf_model_oz_split_ms(bcbf_model* m);

#ifdef __cplusplus
}
#endif
#endif /* BCBF_H_ */
