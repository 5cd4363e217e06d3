"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, one-shot state broadcast, sharded query ==
unsharded query.  The per-rank model is a CPU stand-in built on the oracle (tests only); on GPUs the same
`ShardedPosterior` drives `MVGPModel` over NCCL (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bayesian_cbf_b200.sharding import (ShardedPosterior, broadcast_state, broadcast_state_packed, gather_shards, pack_lower,
                                        packed_lower_elems, shard_bounds, unpack_lower)
from oracle import mvgp_oracle as O


def test_shard_bounds_cover_and_balance():
    for total in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_lower_round_trip():
    g = torch.Generator().manual_seed(0)
    M = torch.tril(torch.randn(384, 384, generator=g, dtype=torch.float64))
    buf = pack_lower(M)
    assert buf.numel() == packed_lower_elems(384) == 128 * 128 * 6
    out = torch.full((384, 384), 7.0, dtype=torch.float64)            # stale contents must not survive
    assert torch.equal(unpack_lower(buf, out), M)


class _OracleModel:
    """CPU stand-in with MVGPModel's interface (fit / alloc_state / state_tensors / query_device)."""

    def __init__(self):
        self.state = None

    def _alloc(self, N, n, p):
        z = lambda *s: torch.zeros(*s, dtype=torch.float64)
        self.state = dict(Linv=z(N, N), alpha=z(N, n), G=z(N, p), W=z(N, n * p), X=z(N, n))

    def fit(self, hyp, X, U, Xdot, jitter, jitter_scale):
        self.hyp = hyp
        N = X.shape[0]
        self._alloc(N, hyp.n, hyp.p)
        UH = O.homogeneous(U)
        Kb = O.gram_train(hyp, X, UH, direct=True) + jitter_scale * torch.diag(jitter)
        L = torch.linalg.cholesky(Kb)
        Linv = torch.linalg.solve_triangular(L, torch.eye(N, dtype=torch.float64), upper=False)
        alpha = Linv.T @ (Linv @ O.residual_targets(hyp, UH, Xdot))
        G = UH @ hyp.B
        for k, v in dict(Linv=Linv, alpha=alpha, G=G, W=(alpha[:, :, None] * G[:, None, :]).reshape(N, -1), X=X).items():
            self.state[k].copy_(v)

    def alloc_state(self, hyp, N):
        self.hyp = hyp
        self._alloc(N, hyp.n, hyp.p)

    def state_tensors(self):
        return self.state

    def query_device(self, Xq, Uq=None, want=('mean', 'svar')):
        s, h = self.state, self.hyp
        Ks = O.rbf_ard(s['X'], Xq, h.lengthscale, h.outputscale, direct=True)
        UHq = O.homogeneous(Uq)
        kb = Ks * (s['G'] @ UHq.T)
        mean = UHq @ h.C + kb.T @ s['alpha']
        v = s['Linv'] @ kb
        svar = h.outputscale * torch.einsum('qa,ab,qb->q', UHq, h.B, UHq) - (v * v).sum(0)
        return dict(mean=mean, svar=svar)


def _problem():
    g = torch.Generator().manual_seed(4)
    f = dict(generator=g, dtype=torch.float64)
    N, n, m, Q = 40, 3, 2, 23
    p = m + 1
    X, U = 2 * torch.rand(N, n, **f) - 1, 2 * torch.rand(N, m, **f) - 1
    Xdot = torch.randn(N, n, **f)
    Ra, Rb = torch.randn(n, n, **f), torch.randn(p, p, **f)
    hyp = O.Hyper(torch.tensor([0.7, 0.9, 1.1], dtype=torch.float64), torch.tensor(1.3, dtype=torch.float64),
                  Ra @ Ra.T + torch.eye(n, dtype=torch.float64), Rb @ Rb.T + torch.eye(p, dtype=torch.float64),
                  0.1 * torch.randn(p, n, **f))
    jit = torch.rand(N, **f)
    Xq, Uq = 2 * torch.rand(Q, n, **f) - 1, 2 * torch.rand(Q, m, **f) - 1
    return hyp, X, U, Xdot, jit, Xq, Uq


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        hyp, X, U, Xdot, jit, Xq, Uq = _problem()
        sp = ShardedPosterior(_OracleModel()).fit(hyp, X, U, Xdot, jit, 1e-5)
        # only the source rank factorised; the others hold the broadcast copy
        ref = _OracleModel()
        ref.fit(hyp, X, U, Xdot, jit, 1e-5)
        for k in ref.state:
            assert torch.equal(sp.model.state[k], ref.state[k]), k
        assert sp.broadcast_bytes == sum(t.numel() * 8 for t in ref.state.values())
        lo, hi, out = sp.query_shard(Xq, Uq)
        assert (lo, hi) == shard_bounds(Xq.shape[0], world, rank)
        full = ref.query_device(Xq, Uq)
        assert torch.equal(out['mean'], full['mean'][lo:hi]) or torch.allclose(out['mean'], full['mean'][lo:hi], rtol=0, atol=1e-14)
        gathered = sp.query_gathered(Xq, Uq)
        assert torch.allclose(gathered['svar'], full['svar'], rtol=0, atol=1e-14)
        assert torch.allclose(gathered['mean'], full['mean'], rtol=0, atol=1e-14)
        # raw helpers
        t = dict(a=torch.full((3,), float(rank)), b=torch.full((2, 2), float(rank)))
        nb = broadcast_state(t, src=1, order=('a', 'b'))
        assert nb == (3 + 4) * 4 and float(t['a'][0]) == 1.0 and float(t['b'][1, 1]) == 1.0
        # the packed one-collective form used for factor-sized states (Npad a multiple of 128)
        g = torch.Generator().manual_seed(9)
        Lref = torch.tril(torch.randn(256, 256, generator=g, dtype=torch.float64))
        others = dict(alpha=torch.randn(256, 4, generator=g, dtype=torch.float64), G=torch.randn(256, 3, generator=g, dtype=torch.float64),
                      W=torch.randn(256, 9, generator=g, dtype=torch.float64), X=torch.randn(200, 3, generator=g, dtype=torch.float64))
        st = dict(Linv=Lref.clone() if rank == 1 else torch.full((256, 256), 3.0, dtype=torch.float64))
        st.update({k: (v.clone() if rank == 1 else torch.zeros_like(v)) for k, v in others.items()})
        nb = broadcast_state_packed(st, src=1)
        assert nb == 8 * (packed_lower_elems(256) + sum(v.numel() for v in others.values()))
        assert torch.equal(st['Linv'], Lref) and all(torch.equal(st[k], v) for k, v in others.items())
        local = torch.arange(*shard_bounds(5, world, rank), dtype=torch.float64).reshape(-1, 1)
        assert torch.equal(gather_shards(local, 5).reshape(-1), torch.arange(5, dtype=torch.float64))
        q.put((rank, 'ok'))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, 'ok'), (1, 'ok')], results
