"""Multi-GPU plumbing for the MVGP path: one process per GPU, `torch.distributed` (NCCL on GPUs; the same code runs over
gloo for the CPU tests).  The path shards by independent units — query states (BASELINE configs[3]) or rollouts
(configs[4]) — with the fitted factor replicated: rank `src` factorises once, the fitted state is broadcast once over
NVLink, and after that there is NO data-path collective.  The reference has no distributed code at all.
"""
import torch
import torch.distributed as dist

# order in which the fitted state is broadcast (MVGPModel.state_tensors keys); L itself is not needed by queries
STATE_ORDER = ('Linv', 'alpha', 'G', 'W', 'X')


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(total, world, rank):
    """Contiguous, balanced partition of range(total): the first `total % world` ranks get one extra unit."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_state(tensors, src=0, group=None, order=STATE_ORDER):
    """In-place broadcast of the fitted-state tensors from `src` (all ranks pass equally-shaped buffers).
    Returns the number of bytes each receiving rank got."""
    nbytes = 0
    if world_info(group)[1] == 1:
        return nbytes
    for key in order:
        t = tensors[key]
        dist.broadcast(t, src=src, group=group)
        nbytes += t.numel() * t.element_size()
    return nbytes


def packed_lower_elems(Npad, block=128):
    nb = Npad // block
    return block * block * nb * (nb + 1) // 2


def pack_lower(M, out=None, block=128):
    """Lower-triangular 128-blocks of a factor-sized matrix (Npad, Npad), block row after block row, into one contiguous
    vector: (nb (nb + 1) / 2) * 128 * 128 elements — what a receiving rank needs of L^-1 (the strictly-upper blocks are
    zero).  One strided copy per block row."""
    Npad = M.shape[0]
    nb = Npad // block
    out = torch.empty(packed_lower_elems(Npad, block), dtype=M.dtype, device=M.device) if out is None else out
    if M.is_cuda and M.dtype is torch.float64 and block == 128 and M.stride(1) == 1:      # one launch (bcbf_pack_lower)
        from . import _lib
        _lib.check(_lib.load().bcbf_pack_lower(M.data_ptr(), M.stride(0), Npad, out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream))
        return out
    off = 0
    for i in range(nb):
        w = (i + 1) * block
        out[off:off + block * w].view(block, w).copy_(M[i * block:(i + 1) * block, :w])
        off += block * w
    return out


def unpack_lower(buf, M, block=128):
    """Inverse of pack_lower into a (Npad, Npad) buffer; the strictly-upper blocks are zeroed."""
    Npad = M.shape[0]
    nb = Npad // block
    if M.is_cuda and M.dtype is torch.float64 and block == 128 and M.stride(1) == 1:
        from . import _lib
        _lib.check(_lib.load().bcbf_unpack_lower(buf.data_ptr(), Npad, M.data_ptr(), M.stride(0),
                                                 torch.cuda.current_stream().cuda_stream))
        return M
    off = 0
    for i in range(nb):
        w = (i + 1) * block
        M[i * block:(i + 1) * block, :w].copy_(buf[off:off + block * w].view(block, w))
        if w < Npad:
            M[i * block:(i + 1) * block, w:].zero_()
        off += block * w
    return M


def broadcast_state_packed(tensors, src=0, group=None, order=STATE_ORDER):
    """ONE collective for the whole fitted state: the packed lower triangle of L^-1 followed by alpha, G, W, X in a
    single contiguous buffer (1.08 GB instead of 2.16 GB in five calls at N = 16384).  In place on the receivers.
    Returns the number of bytes broadcast."""
    rank, world = world_info(group)
    if world == 1:
        return 0
    Linv = tensors['Linv']
    rest = [tensors[k] for k in order if k != 'Linv']
    nl = packed_lower_elems(Linv.shape[0])
    total = nl + sum(t.numel() for t in rest)
    buf = torch.empty(total, dtype=Linv.dtype, device=Linv.device)
    if rank == src:
        pack_lower(Linv, buf[:nl])
        off = nl
        for t in rest:
            buf[off:off + t.numel()].copy_(t.reshape(-1))
            off += t.numel()
    dist.broadcast(buf, src=src, group=group)
    if rank != src:
        unpack_lower(buf[:nl], Linv)
        off = nl
        for t in rest:
            t.copy_(buf[off:off + t.numel()].view(t.shape))
            off += t.numel()
    return total * buf.element_size()


def gather_shards(local, total, group=None):
    """All-gather of per-rank result shards (leading dimension partitioned by shard_bounds) into the full result on
    every rank.  Optional: the bench keeps results sharded."""
    rank, world = world_info(group)
    if world == 1:
        return local
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((width,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


class ShardedPosterior:
    """Factor on one rank, broadcast once, answer query shards everywhere.

    `model` is an `MVGPModel` (or anything with fit / alloc_state / state_tensors / query_device).
    """

    def __init__(self, model, group=None, src=0):
        self.model = model
        self.group = group
        self.src = src
        self.rank, self.world = world_info(group)
        self.broadcast_bytes = 0

    def fit(self, hyper, X, U, Xdot, jitter=None, jitter_scale=1e-5):
        if self.rank == self.src:
            self.model.fit(hyper, X, U, Xdot, jitter, jitter_scale)
        else:
            self.model.alloc_state(hyper, X.shape[0])
        st = self.model.state_tensors()
        packed = st['Linv'].ndim == 2 and st['Linv'].shape[0] == st['Linv'].shape[1] and st['Linv'].shape[0] % 128 == 0
        self.broadcast_bytes = (broadcast_state_packed if packed else broadcast_state)(st, self.src, self.group)
        if self.rank != self.src and hasattr(self.model, 'adopt_state'):
            self.model.adopt_state()
        return self

    def query_shard(self, Xq_all, Uq_all=None, want=('mean', 'svar')):
        """This rank's slice of a replicated query set: returns (lo, hi, outputs)."""
        lo, hi = shard_bounds(Xq_all.shape[0], self.world, self.rank)
        out = self.model.query_device(Xq_all[lo:hi].contiguous(),
                                      None if Uq_all is None else Uq_all[lo:hi].contiguous(), want=want)
        return lo, hi, out

    def query_gathered(self, Xq_all, Uq_all=None, want=('mean', 'svar')):
        lo, hi, out = self.query_shard(Xq_all, Uq_all, want)
        return {k: gather_shards(v, Xq_all.shape[0], self.group) for k, v in out.items()}
