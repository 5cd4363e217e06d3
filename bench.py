#!/usr/bin/env python
"""bench.py — posterior queries/sec (mean + cov of F(x)u) and fit time at N training points.

Workload (BASELINE.json configs[3]): synthetic unicycle MVGP, n=3, m=2, N=16384 training
points, queries in steps of `--queries-per-step` states (default 18944 = 148 SMs x 32 queries x 4 waves;
53 steps = 1.004M queries).  One "step" = one pass of the hot path over one batch of queries: cross-Gram,
posterior mean M_k (3x3), posterior covariance B_k (3x3) [N^2 p flops per query on the FP64 tensor pipe],
u-contraction to mean (3) and scalar variance.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, this repo)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for every key.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DIM, M_DIM = 3, 2
P_DIM = M_DIM + 1


# --------------------------------------------------------------------------------------------- workload
def make_workload(N, seed=0):
    """SURVEY §8d config 4.  Everything is drawn from one CPU generator so CPU and GPU arms see identical bits."""
    g = torch.Generator().manual_seed(seed)
    f64 = dict(dtype=torch.float64, generator=g)
    X = 4 * torch.rand(N, N_DIM, **f64) - 2
    U = 2 * torch.rand(N, M_DIM, **f64) - 1

    def ackermann_F(X, L):  # reference unicycle_move_to_pose.py:235-260  F = [f | g], f = 0
        th = X[:, 2]
        F = torch.zeros(X.shape[0], 3, 3, dtype=torch.float64)
        F[:, 0, 1] = th.cos()
        F[:, 1, 1] = th.sin()
        F[:, 2, 2] = 1.0 / L
        return F
    Ftrue = ackermann_F(X, 1.0) - ackermann_F(X, 12.0)
    UH = torch.cat([torch.ones(N, 1, dtype=torch.float64), U], dim=1)
    Xdot = torch.einsum('inp,ip->in', Ftrue, UH) + 0.01 * torch.randn(N, N_DIM, **f64)
    Ra = torch.randn(N_DIM, N_DIM, **f64)
    Rb = torch.randn(P_DIM, P_DIM, **f64)
    hyp = dict(lengthscale=torch.tensor([0.7, 0.9, 1.1], dtype=torch.float64),
               outputscale=torch.tensor(1.3, dtype=torch.float64),
               A=Ra @ Ra.T + torch.eye(N_DIM, dtype=torch.float64),
               B=Rb @ Rb.T + torch.eye(P_DIM, dtype=torch.float64),
               C=torch.zeros(P_DIM, N_DIM, dtype=torch.float64))
    jitter = torch.rand(N, **f64)
    return X, U, Xdot, hyp, jitter


def make_queries(Q, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    Xq = 4 * torch.rand(Q, N_DIM, dtype=torch.float64, generator=g) - 2
    Uq = 2 * torch.rand(Q, M_DIM, dtype=torch.float64, generator=g) - 1
    return Xq, Uq


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(',')]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                     parts[3:7]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------- CPU legs
def oracle_hyper(hyp):
    from oracle import mvgp_oracle as O
    return O.Hyper(hyp['lengthscale'], hyp['outputscale'], hyp['A'], hyp['B'], hyp['C'])


def cpu_fit(hyp, X, U, Xdot, jitter):
    """Gram + jittered Cholesky (control_affine_model.py:366-377, 899-921) on the host cores; returns (L, seconds)."""
    from oracle import mvgp_oracle as O
    h = oracle_hyper(hyp)
    t0 = time.perf_counter()
    L = O.perturbed_cholesky(h, X, O.homogeneous(U), [jitter])
    return L, time.perf_counter() - t0


def cpu_queries(hyp, X, U, Xdot, L, Xq, Uq, chunk=1024):
    from oracle import mvgp_oracle as O
    h = oracle_hyper(hyp)
    t0 = time.perf_counter()
    out = O.posterior_blocks(h, X, U, Xdot, L, Xq, Uq, chunk=chunk)
    return out, time.perf_counter() - t0


def run_reference_arm(args):
    """The reference's algorithm (restated op for op in oracle/; the reference itself is pure Python over a
    gpytorch fork that cannot be installed here) on the host cores, all threads, same workload config."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    N = args.n_train
    X, U, Xdot, hyp, jitter = make_workload(N)
    L, fit_s = cpu_fit(hyp, X, U, Xdot, jitter)
    qs = args.ref_queries_per_step
    times = []
    for s in range(args.warmup + args.steps):
        Xq, Uq = make_queries(qs, s)
        _, dt = cpu_queries(hyp, X, U, Xdot, L, Xq, Uq, chunk=qs)
        if s >= args.warmup:
            times.append(dt)
    total = float(sum(times))
    value = qs * len(times) / total
    lit_q = max(2, min(args.ref_literal_queries, 64))
    lit_v, _ = cpu_literal_series(hyp, X, U, Xdot, L, make_queries(lit_q, 99)[0], chunk=max(1, lit_q // 2))
    sample = ("N=%d factor built once on the host (%.2f s: Gram + jittered Cholesky), then %d steps of %d queries "
              "(multi-RHS triangular solve + per-query 3x3 blocks, float64)" % (N, fit_s, len(times), qs))
    line = dict(impl='reference', metric='posterior queries/sec (mean+cov of F(x)u) at N train pts', value=value,
                unit='queries/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * total / len(times), higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f64', data='synthetic',
                config=dict(workload='synthetic unicycle MVGP fit N=%d + batched posterior query (BASELINE configs[3])' % N,
                            n_train=N, n=N_DIM, m=M_DIM, queries_per_step=qs,
                            note='bounded CPU sample of the same workload; inputs larger than L2/LLC'),
                fit_s=fit_s,
                cpu_baseline=dict(value=value, unit='queries/s', cores=cores, kind='port', sample=sample,
                                  series='(ii) reference-restated: one multi-RHS triangular solve, per-query blocks',
                                  series_i_reference_literal=dict(
                                      value=lit_v, unit='queries/s',
                                      sample='%d queries, torch.cholesky_solve(kb* (b,N,p), L) with L broadcast over b '
                                             '(control_affine_model.py:1053), chunks of %d' % (lit_q, max(1, lit_q // 2)))),
                e2e=dict(value=value, unit='queries/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- our arm
def measure_dgemm_peak(device):
    """cuBLAS DGEMM 8192^3 on this GPU: the FP64 tensor-pipe roofline denominator (MEASURED_PEAKS.json has no
    FP64 entry).  Library call, used ONLY as the denominator."""
    n = 8192
    a = torch.randn(n, n, device=device, dtype=torch.float64)
    b = torch.randn(n, n, device=device, dtype=torch.float64)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2 * n ** 3 / best * 1e-9


def measure_int8_peaks(gpu_index):
    """The int8 tensor-pipe denominators, measured live on this GPU by tools/microbench/umma_i8_probe.cu (bare tcgen05.mma
    kind::i8 loops out of shared memory, no loads): the issue peak of the pipe (7 MMAs of N=256 per K step, 896 cycles:
    the tensor floor), and what that loop and the kernel's own 10-MMA pattern sustain for ~2 s with RANDOM operand digits
    under the power cap.  MEASURED_PEAKS.json carries no int8 entry.  Returns None if the probe is unavailable."""
    import re
    exe = os.path.join(ROOT, 'build', 'umma_i8_probe')
    if not os.path.exists(exe):
        try:
            import __graft_entry__ as ge
            ge.build()
        except Exception:
            return None
    try:
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=str(gpu_index)) if 'CUDA_VISIBLE_DEVICES' not in os.environ else None
        txt = subprocess.run([exe], capture_output=True, text=True, timeout=180, env=env).stdout
    except Exception:
        return None
    def grab(pat):
        m = re.search(pat, txt)
        return float(m.group(1)) if m else None
    out = dict(issue_peak_tops=grab(r'rate grid=148 7 MMAs of N=256\s*:.*?, ([\d.]+) int8 TOPS'),
               pattern_issue_tops=grab(r'rate grid=148 10-MMA concatenated pattern\s*:.*?, ([\d.]+) int8 TOPS'),
               sustained_random_tops=grab(r'sustained 7 MMAs of N=256\s+random digits:.*?, ([\d.]+) int8 TOPS'),
               pattern_sustained_random_tops=grab(r'sustained 10-MMA concatenated pattern\s+random digits:.*?, ([\d.]+) int8 TOPS'),
               sustained_random_sm_mhz=grab(r'sustained 7 MMAs of N=256\s+random digits:.*?mean SM clock (\d+) MHz'))
    return out if out['issue_peak_tops'] and out['sustained_random_tops'] else None


def cpu_literal_series(hyp, X, U, Xdot, L, Xq, chunk):
    """SURVEY 8d series (i), the reference's statement LITERALLY (control_affine_model.py:1051-1055): kb* (b, N, p),
    `torch.cholesky_solve(kb*, L)` with L broadcast over the b queries (b separate two-sided solves, 8 b N^2 bytes), then the
    mean and the per-query diagonal block of :1079-1088.  Chunked so that the broadcast factor fits in memory."""
    from oracle import mvgp_oracle as O
    h = oracle_hyper(hyp)
    UH = O.homogeneous(U)
    G, Y = UH @ h.B, O.residual_targets(h, UH, Xdot)
    t0 = time.perf_counter()
    for s0 in range(0, Xq.shape[0], chunk):
        xq = Xq[s0:s0 + chunk]
        kb_star = O.rbf_ard(xq, X, h.lengthscale, h.outputscale).unsqueeze(-1) * G.unsqueeze(0)       # (b, N, p)  :1051
        Bdagger = torch.cholesky_solve(kb_star, L)                                                    # :1053
        mean_k = h.C.t().unsqueeze(0) + torch.matmul(Y.t().unsqueeze(0), Bdagger)                      # :1055
        Bk = h.outputscale * h.B.unsqueeze(0) - torch.matmul(kb_star.transpose(-2, -1), Bdagger)       # diagonal blocks of :1079-1088
    return Xq.shape[0] / (time.perf_counter() - t0), (mean_k, Bk)


def run_ours(args):
    import torch.distributed as dist
    from bayesian_cbf_b200 import _lib
    from bayesian_cbf_b200.model import MVGPModel, make_hyper
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        # NCCL prints its version banner with printf when NCCL_DEBUG=VERSION is in the environment: keep file
        # descriptor 1 pointed at stderr while the communicator is created, so that stdout carries the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    lib = _lib.load()
    if args.oz_group:
        _lib.check(lib.bcbf_oz_set_group(args.oz_group))
    N, QS, K, W = args.n_train, args.queries_per_step, args.steps, args.warmup
    X, U, Xdot, hyp, jitter = make_workload(N)
    hyper = make_hyper(N_DIM, P_DIM, hyp['lengthscale'].numpy(), float(hyp['outputscale']), hyp['A'].numpy(),
                       hyp['B'].numpy(), hyp['C'].numpy())
    model = MVGPModel(local_rank).set_var_path(args.var_path).set_oz_digits(args.digits)
    nprod = {7: 28.0, 6: 21.0}[args.digits]
    i8 = args.var_path == 'int8'
    prof_enable = lib.bcbf_oz_profile_enable if i8 else lib.bcbf_profile_enable
    prof_read = lib.bcbf_oz_profile_read if i8 else lib.bcbf_profile_read

    # ---- fit (Gram + jittered Cholesky + L^-1 + alpha): rank 0 factorises, NCCL broadcasts the state ----------
    fit, fit_cold = dict(), dict()
    strong = bool(args.strong)
    if rank == 0:
        # first fit: cold (allocations, clock ramp-up, lazy module loading) — reported as fit_cold_ms, never as fit_ms
        model.fit(hyper, X.numpy(), U.numpy(), Xdot.numpy(), jitter.numpy(), 1e-5)
        fit_cold = model.fit_timing_ms()
    else:
        model.alloc_state(hyper, N)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if rank == 0:
        model.fit(hyper, X.numpy(), U.numpy(), Xdot.numpy(), jitter.numpy(), 1e-5)
        fit = model.fit_timing_ms()
    bcast_ms, bcast_bytes = 0.0, 0
    if world > 1:
        st = model.state_tensors()
        warm = torch.empty(64 << 20, dtype=torch.uint8, device=dev)     # NCCL sets up its large-message protocol / buffer
        dist.broadcast(warm, src=0)                                      # registration on first use: not part of the state
        del warm
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        from bayesian_cbf_b200.sharding import broadcast_state_packed
        bcast_bytes = broadcast_state_packed(st, src=0)   # packed lower triangle of L^-1 + alpha, G, W, X: ONE collective,
                                                          # once; no collective during querying
        if rank != 0:
            model.adopt_state()
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    fit_wall_s = time.perf_counter() - t0

    # ---- queries: weak scaling, every rank runs K steps of QS queries on its own shard -------------------------
    K_total = K
    if strong:
        lo_, hi_ = (rank * K) // world, ((rank + 1) * K) // world      # this rank's share of the K batches
        K = hi_ - lo_
    steps_total = W + K
    Xq_all, Uq_all = make_queries(QS * steps_total, 7919 * rank)
    Xq_d, Uq_d = Xq_all.to(dev), Uq_all.to(dev)   # resident in HBM before the timed region
    outs = None

    def step(i):
        s = slice(i * QS, (i + 1) * QS)
        return model.query_device(Xq_d[s], Uq_d[s])

    for i in range(W):
        outs = step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    prof_enable(1)
    launches0 = lib.bcbf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    # cudaProfilerStart/Stop bracket exactly the timed region: `ncu --profile-from-start off` then lists its launches
    # only (profiles/*_ncu_launches_bench.csv); without a profiler attached the two calls do nothing
    torch.cuda.cudart().cudaProfilerStart()
    e0.record()
    for i in range(W, W + K):
        outs = step(i)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.bcbf_launch_count() - launches0
    import ctypes
    kms, kn = ctypes.c_double(), ctypes.c_int()
    prof_read(ctypes.byref(kms), ctypes.byref(kn))
    prof_enable(0)
    clocks = sampler.stop() if rank == 0 else {}
    t = torch.tensor([ms, fit.get('total', 0.0) or 0.0, bcast_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, fit_ms_all, bcast_ms_all = float(t[0]), float(t[1]), float(t[2])
    checksum = float(outs['svar'].sum().item())

    # ---- e2e: same metric through the C ABI with HOST buffers (pinned), copies inside the timed region ----------
    pin = lambda *shape: torch.empty(*shape, dtype=torch.float64).pin_memory()
    hXq, hUq = pin(QS, N_DIM), pin(QS, M_DIM)
    hmean, hsvar, hMk, hBk = pin(QS, N_DIM), pin(QS), pin(QS, N_DIM, P_DIM), pin(QS, P_DIM, P_DIM)
    Ke = max(2, min(K, args.e2e_steps))
    e2e_times = []
    for i in range(W + Ke):
        s = slice((i % steps_total) * QS, (i % steps_total + 1) * QS)
        hXq.copy_(Xq_all[s])
        hUq.copy_(Uq_all[s])
        if world > 1:
            dist.barrier()
        t1 = time.perf_counter()
        model.query_into(hXq.numpy(), hUq.numpy(), hmean.numpy(), hsvar.numpy(), hMk.numpy(), hBk.numpy())
        dt = time.perf_counter() - t1      # bcbf_model_query is synchronous (stream sync inside)
        if i >= W:
            e2e_times.append(dt)
    te = torch.tensor([sum(e2e_times)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * QS * len(e2e_times) / float(te.item())
    h2d = QS * (N_DIM + M_DIM) * 8
    d2h = QS * (N_DIM + 1 + N_DIM * P_DIM + P_DIM * P_DIM) * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------
    kernel_ms = kms.value / max(kn.value, 1)
    fp64_flops_per_launch = float(N) * N * P_DIM * QS          # N^2 p FP64 flops per query
    peak_dgemm = measure_dgemm_peak(dev)
    if i8:
        # oz_var_kernel runs the contraction on the int8 tensor pipe: 7 x 7 digit products with digit sum <= 6 = 28
        # int8 GEMMs of the same (triangular) shape, N^2/2 MACs per column each -> 28 N^2 p int8 ops per query
        ops_per_launch = nprod * fp64_flops_per_launch
        achieved = ops_per_launch / (kernel_ms * 1e-3) * 1e-12
        # denominator: the int8 tensor pipe measured LIVE on this GPU (MEASURED_PEAKS.json has HBM and bf16 only).
        # The kernel runs inside a seconds-long step under the power cap, so `peak` is the sustained figure: what the bare
        # tensor-floor MMA loop holds for ~2 s on random operand digits; the burst issue peak is given beside it.
        i8p = measure_int8_peaks(local_rank)
        bf16x2 = None
        mp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        if os.path.exists(mp):
            try:
                bf16x2 = 2.0 * float(json.load(open(mp))['bf16_tflops_sustained'])
            except Exception:
                bf16x2 = None
        if i8p is not None:
            peak = i8p['sustained_random_tops']
            peak_source = ('tcgen05.mma kind::i8 loop of tools/microbench/umma_i8_probe.cu measured live on this GPU: '
                           'sustained ~2 s, random digits, power cap active (7 MMAs of N=256 per K step = the 896-cycle '
                           'tensor floor); of measured')
        elif bf16x2 is not None:
            peak = bf16x2
            peak_source = ('2 x bf16_tflops_sustained of MEASURED_PEAKS.json (int8 probe unavailable), of measured')
        else:
            peak = 2.0 * 1400.0
            peak_source = '2 x 1.4 PFLOP/s sustained bf16 (B200_PROFILING.md fallback), of fallback'
        tname, kname = 'oz_var_ncu_summary.json', 'oz_var_kernel'
        extra = dict(unit_note='int8 tensor ops (TOP/s)', algorithmic_ops_per_launch=ops_per_launch,
                     fp64_equivalent_tflops=fp64_flops_per_launch / (kernel_ms * 1e-3) * 1e-12,
                     cublas_dgemm_tflops_live=peak_dgemm, int8_probe=i8p,
                     frac_of_issue_peak=(achieved / i8p['issue_peak_tops']) if i8p else None,
                     frac_of_own_pattern_sustained=(achieved / i8p['pattern_sustained_random_tops'])
                     if i8p and i8p.get('pattern_sustained_random_tops') else None,
                     frac_of_2x_bf16_sustained=(achieved / bf16x2) if bf16x2 else None)
    else:
        achieved = fp64_flops_per_launch / (kernel_ms * 1e-3) * 1e-12
        peak = peak_dgemm
        peak_source = ('cuBLAS DGEMM 8192^3 measured live on this GPU (FP64 tensor pipe; MEASURED_PEAKS.json has no FP64 '
                       'entry; DMMA instruction peak measured 37.1 TFLOP/s, profiles/r01_fp64_peaks.txt)')
        tname, kname = 'post_var_ncu_summary.json', 'post_var_kernel'
        extra = dict(algorithmic_flops_per_launch=fp64_flops_per_launch)
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', tname)
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get('dram_bytes_per_launch')
        except Exception:
            traffic = None
    roofline = dict(bound='tensor', achieved=achieved, peak=peak, unit='TFLOP/s', frac=achieved / peak,
                    traffic=traffic, kernel=kname, kernel_ms=kernel_ms, kernel_share_of_step=kms.value / ms,
                    peak_source=peak_source, **extra)

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 only, N=1 only --------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        Lc, cfit_s = cpu_fit(hyp, X, U, Xdot, jitter)
        qs = args.cpu_sample_queries
        (cMk, cBk, cmean, csvar), cq_s = cpu_queries(hyp, X, U, Xdot, Lc, Xq_all[:qs], Uq_all[:qs], chunk=qs)
        # parity spot-check of the benchmarked path against the oracle on the sample
        got = model.query(Xq_all[:qs].numpy(), Uq_all[:qs].numpy())
        prior = float(hyp['outputscale'] * torch.linalg.matrix_norm(hyp['B'], 2))
        parity = dict(Bk_rel=float(np.abs(got['Bk'] - cBk.numpy()).max() / prior),
                      svar_rel=float(np.abs(got['svar'] - csvar.numpy()).max() / prior),
                      mean_rel=float(np.abs(got['mean'] - cmean.numpy()).max() / max(1e-300, np.abs(cmean.numpy()).max())),
                      mean_rel_note='against the reference-form LAPACK evaluation (gpytorch expanded-distance Gram, '
                                    'cholesky_solve)')
        if not args.no_parity_floor:
            # the mean against the EXACT solution of the reference's linear system, next to the measured floor of the
            # reference's own float64 arithmetic (oracle.mean_parity_floor; tests/parity_util.py explains the numbers)
            from tests.parity_util import mean_reference
            del Lc
            t0 = time.perf_counter()
            ref = mean_reference(oracle_hyper(hyp), X, U, Xdot, jitter, Xq_all[:qs], Uq_all[:qs])
            sc = float(ref['mean_exact'].abs().max())
            parity.update(mean_rel_exact=float(np.abs(got['mean'] - ref['mean_exact'].numpy()).max() / sc),
                          Mk_rel_exact=float(np.abs(got['Mk'] - ref['Mk_exact'].numpy()).max()
                                             / float(ref['Mk_exact'].abs().max())),
                          floor_ulp_sensitivity=ref['ulp_sensitivity'], floor_lapack_vs_exact=ref['lapack_vs_exact'],
                          tol_exact=ref['tol_exact'], tol_lapack=ref['tol_lapack'],
                          floor_seconds=time.perf_counter() - t0,
                          within_tolerance=bool(parity['Bk_rel'] < 1e-9 and parity['svar_rel'] < 1e-9))
            parity['within_tolerance'] = bool(parity['within_tolerance'] and parity['mean_rel_exact'] < ref['tol_exact']
                                              and parity['mean_rel'] < ref['tol_lapack'])
        cpu = dict(value=qs / cq_s, unit='queries/s', cores=cores, kind='port',
                   sample='N=%d: host Gram + Cholesky once (%.2f s), then %d queries in %.2f s (multi-RHS triangular '
                          'solve + per-query blocks, torch float64, %d threads)' % (N, cfit_s, qs, cq_s, cores),
                   fit_s=cfit_s, parity_vs_gpu=parity)

    line = dict(metric='posterior queries/sec (mean+cov of F(x)u) at N train pts', value=world * QS * K / (ms_max * 1e-3),
                unit='queries/s', n_gpus=world, steps=K, warmup=W, ms_per_step=ms_max / K, higher_is_better=True,
                scaling='strong' if strong else 'weak', vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload='synthetic unicycle MVGP fit N=%d + batched posterior query, %d queries/step/GPU '
                                     '(BASELINE configs[3]; 53 steps = 1.004M queries)' % (N, QS),
                            n_train=N, n=N_DIM, m=M_DIM, queries_per_step=QS, outputs='M_k(3x3), B_k(3x3), mean(3), svar',
                            covariance_kernel=('oz_var_kernel: tcgen05 int8 tensor cores, %dx%d error-free digit splitting of '
                                               'both FP64 operands (%d digit products), FP64 recombination'
                                               % (args.digits, args.digits, int(nprod)) if i8 else
                                               'post_var_kernel: FP64 tensor pipe (DMMA)'),
                            parallelism='queries sharded, factor broadcast once (NCCL)' if world > 1 else 'single GPU',
                            l2='inputs larger than L2: L^-1 is %.2f GB (lower triangle), streamed every step' % (4.0 * N * (N + 1) / 1e9)),
                fit_ms=fit.get('total'), fit_breakdown_ms=fit, fit_cold_ms=fit_cold.get('total'),
                fit_note='fit_ms: second fit of this process (the first, fit_cold_ms, pays allocations and clock ramp-up)',
                fit_wall_s=fit_wall_s, factor_broadcast_ms=bcast_ms,
                factor_broadcast_gbs=(bcast_bytes / (bcast_ms * 1e-3) * 1e-9) if bcast_ms > 0 else None,
                clocks=dict(sm_mhz=clocks.get('sm_mhz'), sm_max_mhz=clocks.get('sm_max_mhz'), reasons=clocks.get('reasons', []),
                            samples=clocks.get('samples', 0)),
                e2e=dict(value=e2e_value, unit='queries/s', h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=len(e2e_times)),
                gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu, checksum_svar=checksum)
    if strong:
        # strong scaling: K_total batches divided over the ranks; time = warm fit on rank 0 + factor broadcast + the slowest
        # rank's query time, the three device-timed segments summed (input synthesis and warm-up steps are outside)
        tot_ms = fit_ms_all + bcast_ms_all + ms_max
        line.update(value=QS * K_total / (tot_ms * 1e-3), steps=K_total, ms_per_step=tot_ms / K_total,
                    strong=dict(total_queries=QS * K_total, fit_ms=fit_ms_all, broadcast_ms=bcast_ms_all, query_ms=ms_max,
                                steps_this_rank=K))
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- rollouts (configs[4])
def cpu_rollout_sample(R, N, steps, seed=0):
    """Host restatement of ONE control step of a learning rollout, per rollout: posterior blocks of its own N-point MVGP
    (oracle.posterior_blocks), closed-form CBC / CLC cone terms and the barrier SOCP (tests/fake_ops.py,
    oracle/socp_oracle.py).  Returns rollout-steps per second over R rollouts x `steps` steps (refits excluded)."""
    import math
    from oracle import mvgp_oracle as O
    from tests import fake_ops
    from bayesian_cbf_b200 import unicycle as Un

    class _P:
        def setattr(self, o, n, v):
            setattr(o, n, v)
    fake_ops.installed(_P()).__enter__()
    g = torch.Generator().manual_seed(seed)
    f = dict(dtype=torch.float64, generator=g)
    models = []
    eye = torch.eye(3, dtype=torch.float64)
    for r in range(R):
        X = torch.zeros(N, 3, dtype=torch.float64)
        X[:, 2] = 2 * torch.rand(N, **f) - 1
        U = 2 * torch.rand(N, 2, **f) - 1
        Xdot = 0.1 * torch.randn(N, 3, **f)
        hyp = O.Hyper(torch.tensor([1.0, 1.0, 0.7], dtype=torch.float64), torch.tensor(1.0, dtype=torch.float64), eye, eye,
                      torch.zeros(3, 3, dtype=torch.float64))
        L = O.perturbed_cholesky(hyp, X, O.homogeneous(U), [torch.rand(N, **f)], direct=True)
        models.append((hyp, X, U, Xdot, L))
    x0, xg = [-3.0, -1.0, -math.pi / 4], [0.0, 0.0, math.pi / 4]
    planner = Un.PiecewiseLinearPlanner(x0, xg, 2000, 0.001, frac_time_to_reach_goal=0.95)
    cbfs = Un.obstacles_at_mid_from_start_and_goal(x0, xg, term_weights=(0.7, 0.3))

    def posterior(Xs):
        Mk, Bk = [], []
        for r, (hyp, X, U, Xdot, L) in enumerate(models):
            m, b = O.posterior_blocks(hyp, X, U, Xdot, L, Xs[r:r + 1], direct=True)
            Mk.append(m[0]); Bk.append(b[0])
        return torch.stack(Mk), torch.stack(Bk), eye
    ctrl = Un.BayesCBFController(planner, Un.CLFCartesian(Kp=(0.9, 1.5, 0.0)), cbfs, [5.0, 5.0], model_L=12.0,
                                 clf_gamma=10.0, max_risk=0.01, posterior=posterior)
    X0 = torch.tensor(x0, dtype=torch.float64).repeat(R, 1) + 0.05 * (torch.rand(R, 3, **f) - 0.5)
    Un.rollout(ctrl, X0, 2, 0.001, true_L=1.0)
    t0 = time.perf_counter()
    Un.rollout(ctrl, X0, steps, 0.001, true_L=1.0)
    return R * steps / (time.perf_counter() - t0)


def run_rollouts(args):
    """BASELINE configs[4] shape on this rank's GPU: R independent `learning_helps_avoid_getting_stuck` rollouts (reference
    unicycle_move_to_pose.py:1948-1969), each with its own MVGP refitted every `--train-every` steps on at most 200 of its
    own samples with `--adam-iters` Adam steps; per control step and rollout: ensemble posterior -> CBC / CLC cone terms
    -> batched SOCP, the control step replayed from a CUDA graph.  One "step" = one control step of all R rollouts."""
    import math
    import torch.distributed as dist
    from bayesian_cbf_b200 import _lib, unicycle as Un
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        Rc, Sc = 4, max(10, min(args.steps, 40))
        v = cpu_rollout_sample(Rc, 200, Sc)
        sample = '%d rollouts x %d control steps, N=200 per-rollout models, host restatement (refits excluded)' % (Rc, Sc)
        print(json.dumps(dict(impl='reference', metric='controlled rollout steps/sec (posterior + CBC terms + SOCP per step)',
                              value=v, unit='rollout-steps/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                              ms_per_step=1e3 * Rc / v, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
                              data='synthetic', config=dict(workload='ensemble of unicycle learning rollouts (BASELINE configs[4])'),
                              cpu_baseline=dict(value=v, unit='rollout-steps/s', cores=cores, kind='port', sample=sample),
                              e2e=dict(value=v, unit='rollout-steps/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    lib = _lib.load()
    R, K, W, dt = args.rollouts, args.steps, args.warmup, 0.001
    x0, xg = [-3.0, -1.0, -math.pi / 4], [0.0, 0.0, math.pi / 4]
    planner = Un.PiecewiseLinearPlanner(x0, xg, 2000, dt, frac_time_to_reach_goal=0.95)
    cbfs = Un.obstacles_at_mid_from_start_and_goal(x0, xg, term_weights=(0.7, 0.3))
    learner = Un.EnsembleLearner(R, dt, model_L=12.0, max_train=200, train_every_n_steps=args.train_every,
                                 lengthscale=(1.0, 1.0, 0.7), outputscale=1.0, adam_iters=args.adam_iters, seed=rank,
                                 device=dev, query_raw_state=not args.query_shift_invariant)
    ctrl = Un.BayesCBFController(planner, Un.CLFCartesian(Kp=(0.9, 1.5, 0.0)), cbfs, [5.0, 5.0], model_L=12.0,
                                 clf_gamma=10.0, max_risk=0.01, posterior=learner.posterior)
    g = torch.Generator().manual_seed(rank)
    hX0 = (torch.tensor(x0, dtype=torch.float64).repeat(R, 1)
           + 0.05 * (torch.rand(R, 3, generator=g, dtype=torch.float64) - 0.5)).pin_memory()
    X0 = hX0.to(dev)
    Un.rollout(ctrl, X0, max(W, 3), dt, true_L=1.0)          # warm-up steps (eager), nothing recorded
    learner.Xs, learner.Us = [], []
    gr = Un.GraphedRollout(ctrl, X0, dt, true_L=1.0).capture()
    state = dict(refits=0)

    def on_step(t, X, u, xdot, ok):
        learner.record(t, X, u, xdot, ok)
        if learner.refits != state['refits']:      # the posterior switched kernels / buffers: re-capture the step
            state['refits'] = learner.refits
            gr.capture()
            gr._set_plan(t)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = lib.bcbf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    X0.copy_(hX0, non_blocking=True)                           # e2e: start states come from pinned host memory ...
    gr.X.copy_(X0)
    out = gr.run(K, on_step=on_step, record=False)
    hXf = gr.X.to('cpu', non_blocking=True)                    # ... and the final states / alive flags go back
    halive = out['alive'].to('cpu', non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = lib.bcbf_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else {}
    t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
    al = torch.tensor([int(halive.sum())], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(al, op=dist.ReduceOp.SUM)
    ms_max, wall_ms = float(t[0]), float(t[1])
    # roofline of the HBM-bound kernel of the step: one rollout's own L^-1 streamed per rollout and step
    roof = None
    if learner.fitted:
        Xq = gr.X.clone()
        for _ in range(3):
            learner.ens.posterior(Xq)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        tk = 0.0
        for _ in range(5):
            flush.zero_()                                      # > L2: every launch streams its factors from HBM
            k0.record()
            learner.ens.posterior(Xq)
            k1.record()
            torch.cuda.synchronize()
            tk += k0.elapsed_time(k1) / 5
        hbm = 6541.1
        mp = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        src = 'fallback'
        if os.path.exists(mp):
            try:
                hbm, src = float(json.load(open(mp))['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs, of measured'
            except Exception:
                pass
        ach = learner.ens.posterior_bytes() / (tk * 1e-3) * 1e-9
        roof = dict(bound='hbm', achieved=ach, peak=hbm, unit='GB/s', frac=ach / hbm, traffic=None,
                    kernel='ens_posterior_kernel', kernel_ms=tk, peak_source=src,
                    algorithmic_bytes_per_launch=learner.ens.posterior_bytes(),
                    note='timed alone after an L2 flush; inside the captured step it is %.0f%% of the step by this time'
                         % (100 * tk / (ms_max / K)))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        v = cpu_rollout_sample(4, 200, 25)
        cpu = dict(value=v, unit='rollout-steps/s', cores=cores, kind='port',
                   sample='4 rollouts x 25 control steps, N=200 per-rollout models, host restatement of posterior + cone '
                          'terms + SOCP (refits excluded)')
    Rt = R * world
    line = dict(metric='controlled rollout steps/sec (posterior + CBC terms + SOCP per step)',
                value=Rt * K / (ms_max * 1e-3), unit='rollout-steps/s', n_gpus=world, steps=K, warmup=max(W, 3),
                ms_per_step=ms_max / K, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64', data='synthetic',
                config=dict(workload='ensemble of %d unicycle learning rollouts per GPU (BASELINE configs[4]: '
                                     'learning_helps_avoid_getting_stuck), refit every %d steps on <= 200 samples with %d '
                                     'Adam steps' % (R, args.train_every, args.adam_iters),
                            rollouts_per_gpu=R, parallelism='rollouts sharded, no collective' if world > 1 else 'single GPU',
                            query_state='raw (reference)' if not args.query_shift_invariant else 'shift-invariant',
                            l2='per-rollout factors: %d x %.0f KB per step, larger than L2' %
                               (R, (learner.ens.posterior_bytes() / R / 1024) if learner.fitted else 0)),
                refits=learner.refits, alive_at_end=int(al.item()), n_train_last=getattr(learner.ens, 'N', 0),
                clocks=dict(sm_mhz=clocks.get('sm_mhz'), sm_max_mhz=clocks.get('sm_max_mhz'), reasons=clocks.get('reasons', []),
                            samples=clocks.get('samples', 0)),
                e2e=dict(value=Rt * K / (wall_ms * 1e-3), unit='rollout-steps/s', h2d_bytes_per_step=R * 3 * 8 / K,
                         d2h_bytes_per_step=R * (3 * 8 + 1) / K, note='whole run by the host clock: start states from pinned '
                         'host memory, final states and alive flags back; refits and graph re-captures included'),
                gpu_launches=int(launches), roofline=roof, cpu_baseline=cpu)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--workload', default='posterior', choices=['posterior', 'rollouts'],
                    help='posterior: BASELINE configs[3] (the metric headline); rollouts: configs[4] ensemble of learning rollouts')
    ap.add_argument('--steps', type=int, default=None, help='default 53 (posterior) / 2000 (rollouts)')
    ap.add_argument('--rollouts', type=int, default=512, help='rollouts per GPU (workload rollouts)')
    ap.add_argument('--train-every', type=int, default=400)
    ap.add_argument('--adam-iters', type=int, default=100)
    ap.add_argument('--query-shift-invariant', action='store_true',
                    help='query the learned GP at [0,0,theta] instead of the raw state (the reference queries raw)')
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--n-train', type=int, default=16384)
    ap.add_argument('--queries-per-step', type=int, default=None,
                    help='default: 18944 (dmma: 148 SMs x 32 x 4) / 18648 (int8: 148 SMs x 21 x 6)')
    ap.add_argument('--var-path', default='int8', choices=['dmma', 'int8'],
                    help='kernel of the N^2 p covariance contraction: FP64 tensor pipe, or int8 tensor cores (tcgen05) with '
                         'error-free digit splitting')
    ap.add_argument('--digits', type=int, default=7, choices=[6, 7],
                    help='digits per operand of the int8 covariance kernel: 7 (default, 28 digit products, FP64 rounding '
                         'level) or 6 (opt-in: 21 products, B_k to ~3e-11 of the prior scale)')
    ap.add_argument('--oz-group', type=int, default=None, help='row blocks per scheduling group of oz_var_kernel (tuning)')
    ap.add_argument('--e2e-steps', type=int, default=8)
    ap.add_argument('--cpu-sample-queries', type=int, default=2048)
    ap.add_argument('--ref-queries-per-step', type=int, default=1024)
    ap.add_argument('--ref-literal-queries', type=int, default=8,
                    help='queries timed with the reference-literal broadcast cholesky_solve (series i)')
    ap.add_argument('--strong', action='store_true',
                    help='strong scaling: the --steps batches are divided over the ranks and the clock also covers the fit '
                         'on rank 0 and the factor broadcast')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity-floor', action='store_true',
                    help='skip the exact-solution / floor measurement of the mean parity (about a minute of host time)')
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 53 if args.workload == 'posterior' else 2000
    if args.workload == 'rollouts':
        return run_rollouts(args)
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.queries_per_step is None:
        args.queries_per_step = 18648 if args.var_path == 'int8' else 18944
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
