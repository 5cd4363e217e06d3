"""Drop-in host API of the MVGP regressor — same class / method names, argument meaning, shapes and error behaviour as
the reference's bayes_cbf/control_affine_model.py, with every N-sized computation running in the hand-written CUDA
kernels of libbcbf.so (no CPU path: a regressor placed on a non-CUDA device raises at the first computation).

    reference                                              here
    ---------------------------------------------------    ---------------------------------------------------------
    ControlAffineExactGP (:139-218)                        same name; parameters via gp_modules (gpytorch names)
    ControlAffineRegressor.fit (:268-335)                  Adam + MultiStepLR on the fused GPU log marginal (mll.py)
    _perturbed_cholesky[_compute], make_psd (:366-385,     fused Gram -> blocked DMMA Cholesky with the 10x jitter retry;
        :899-921)                                          L and L^-1 cached under the reference's cache key
    custom_predict (:390-613)                              control-affine cross Gram + L^-1 products (DMMA)
    ControlAffineRegressorExact (:930-1096)                frakB Gram + L^-1 products; per-query block fast path
    closures f_func_* / fu_func_* / covar_fu_f (:685-848)  same names (consumed by gp_algebra / cbc1 / cbc2)

The random Cholesky jitter is drawn with `torch.rand` on the CPU generator (the reference draws it on the tensor's
device, :907-910), so that a seeded CPU run of the reference and a seeded run of this class consume the same
numbers; the jitter vectors can also be supplied explicitly (`set_jitter_source`) for parity tests.

Extension (not in the reference, needed for the 1M-query workload where a (b,b,p,p) result is impossible):
`custom_predict_blocks(X, U=None)` -> per-query M_k (b,n,p), B_k (b,p,p) [, mean (b,n), svar (b,)].
"""
import logging
import warnings
from functools import partial

import numpy as np
import torch
from torch import nn

from . import autograd_ops, ops
from ._lib import NotPositiveDefiniteError
from .gp_algebra import GaussianProcess
from .gp_modules import (ConstantMean, GammaPrior, IndexKernel, LinearKernel, MultivariateNormalResult, RBFKernel,
                         ScaleKernel)
from .matrix_variate_multitask_kernel import (HetergeneousCoregionalizationKernel, HetergeneousMatrixVariateKernel,
                                              MatrixVariateIndexKernel)
from .matrix_variate_multitask_model import HetergeneousMatrixVariateMean
from .misc import DynamicsModel, torch_kron
from .mll import dense_log_marginal, mvgp_log_marginal

LOG = logging.getLogger(__name__)
LOG.setLevel(logging.INFO)


class CatEncoder:
    """Encodes / decodes arrays by concatenation along the last axis (reference :74-100)."""

    def __init__(self, *sizes):
        self.sizes = list(sizes)

    @classmethod
    def from_data(cls, *arrays):
        self = cls(*[A.shape[-1] for A in arrays])
        return self, self.encode(*arrays)

    def encode(self, *arrays):
        if isinstance(arrays[0], torch.Tensor):
            return torch.cat(arrays, dim=-1)
        return np.concatenate(arrays, axis=-1)

    def decode(self, X):
        idxs = np.cumsum([0] + self.sizes)
        return [X[..., s:e] for s, e in zip(idxs[:-1], idxs[1:])]

    def state_dict(self):
        return dict(sizes=self.sizes)

    def load_state_dict(self, state_dict):
        self.sizes = state_dict['sizes']


class IdentityLikelihood(nn.Module):
    """y = f(x) exactly (reference :103-136): `marginal` is the identity and `noise` reads 0."""

    def __init__(self):
        super().__init__()
        self.min_possible_noise = 1e-6

    @property
    def noise(self):
        return 0

    @noise.setter
    def noise(self, _):
        LOG.warning("Ignore setting of noise")

    def marginal(self, function_dist, *params, **kwargs):
        return function_dist

    def forward(self, function_dist, *params, **kwargs):
        return function_dist


class ControlAffineExactGP(nn.Module):
    """Heterogeneous MVGP model (reference :139-218): MXU = [M, X, UH]; M = 1 rows observe F(x)[1;u], M = 0 rows F(x)."""

    def __init__(self, x_dim, u_dim, likelihood, rank=None, gamma_length_scale_prior=None):
        super().__init__()
        self.likelihood = likelihood
        self.matshape = (1 + u_dim, x_dim)
        self.decoder = CatEncoder(1, x_dim, 1 + u_dim)
        self.mean_module = HetergeneousMatrixVariateMean(ConstantMean(), self.decoder, self.matshape)
        self.task_covar = MatrixVariateIndexKernel(
            IndexKernel(num_tasks=self.matshape[1], rank=(self.matshape[1] if rank is None else rank)),
            IndexKernel(num_tasks=self.matshape[0], rank=(self.matshape[0] if rank is None else rank)))
        prior = None if gamma_length_scale_prior is None else GammaPrior(*gamma_length_scale_prior)
        self.input_covar = ScaleKernel(RBFKernel(ard_num_dims=x_dim, lengthscale_prior=prior))
        self.covar_module = HetergeneousMatrixVariateKernel(self.task_covar, self.input_covar, self.decoder)
        self.train_inputs = None
        self.train_targets = None

    def set_train_data(self, Xtrain, Utrain, XdotTrain):
        assert self.matshape == (1 + Utrain.shape[-1], Xtrain.shape[-1])
        assert Xtrain.shape[-1] == XdotTrain.shape[-1]
        _, MXUtrain = self.encode_from_XU(Xtrain, Utrain, 1)
        self.train_inputs = (MXUtrain,)
        self.train_targets = XdotTrain.reshape(-1)

    def encode_from_XU(self, Xtrain, Utrain=None, M=0):
        Mtrain = Xtrain.new_full([Xtrain.size(0), 1], M)
        if M:
            assert Utrain is not None
            UHtrain = torch.cat([Mtrain, Utrain], dim=1)
        else:
            UHtrain = Xtrain.new_zeros((Xtrain.size(0), self.matshape[0]))
        return CatEncoder.from_data(Mtrain, Xtrain, UHtrain)

    def forward(self, mxu):
        return MultivariateNormalResult(self.mean_module(mxu), self.covar_module(mxu))

    def state_dict(self, *a, **k):
        return dict(matshape=self.matshape,
                    decoder=self.decoder.state_dict(),
                    mean_module=self.mean_module.state_dict(),
                    task_covar=nn.Module.state_dict(self.task_covar),
                    input_covar=nn.Module.state_dict(self.input_covar),
                    train_inputs=self.train_inputs,
                    train_targets=self.train_targets)

    def load_state_dict(self, state_dict, *a, **k):
        sd = dict(state_dict)
        self.matshape = sd.pop('matshape')
        self.train_inputs = sd.pop('train_inputs')
        self.train_targets = sd.pop('train_targets')
        self.decoder.load_state_dict(sd['decoder'])
        self.mean_module.load_state_dict(sd['mean_module'])
        self.task_covar.load_state_dict(sd['task_covar'])
        self.input_covar.load_state_dict(sd['input_covar'])
        return self


def default_device():
    return 'cuda' if torch.cuda.is_available() else 'cpu'


def _need_cuda(device):
    if torch.device(device).type != 'cuda':
        raise RuntimeError("bayesian_cbf_b200: the MVGP path runs on a CUDA device only (no CPU fallback); this "
                           "regressor lives on %r" % (device,))


def _draw_jitter(n, dtype):
    """One U(0,1)^n draw from the CPU generator (reference make_psd :907-910 draws `torch.rand` per attempt)."""
    return torch.rand(n, dtype=dtype)


def make_psd(Kb, cholesky_tries=10, cholesky_perturb_init=1e-5, cholesky_perturb_scale=10, jitters=None):
    """Kb + factor * diag(U(0,1)) with the reference's retry schedule (:899-921); returns (Kbp, lower factor).
    Kb is a CUDA tensor (n, n); the factorisation is the blocked DMMA Cholesky.  `jitters`: optional iterator of
    explicit U(0,1)^n vectors (parity tests)."""
    _need_cuda(Kb.device)
    n = Kb.shape[0]
    npad = ops.padded(n)
    factor = cholesky_perturb_init
    K64 = Kb.double()
    for ntry in range(cholesky_tries):
        eps = next(jitters) if jitters is not None else _draw_jitter(n, Kb.dtype)
        eps = eps.to(device=Kb.device, dtype=torch.float64).contiguous()
        buf = torch.eye(npad, dtype=torch.float64, device=Kb.device)
        buf[:n, :n] = K64
        try:
            L, _ = ops.potrf_(buf, n, eps, factor)
            Kbp = K64 + factor * torch.diag(eps)
            return Kbp.to(Kb.dtype), L[:n, :n].to(Kb.dtype)
        except RuntimeError as e:
            if ntry == cholesky_tries - 1:
                raise
            LOG.warning("Cholesky failed with perturb={} on error {}".format(factor, str(e)))
            factor = factor * cholesky_perturb_scale
    raise AssertionError("unreachable")


def is_psd(X):
    try:
        buf = torch.eye(ops.padded(X.shape[0]), dtype=torch.float64, device=X.device)
        buf[:X.shape[0], :X.shape[0]] = X.double()
        ops.potrf_(buf, X.shape[0], None, 0.0)
    except RuntimeError as e:
        print(e)
        return False
    return True


class ControlAffineRegressor(DynamicsModel):
    """Scikit-like wrapper: F(X), COV(F(X)) = ControlAffineRegressor().fit(X, U, Xdot).predict(Xtest)."""
    ground_truth = False

    def __init__(self, x_dim, u_dim, device=None, default_device=default_device, gamma_length_scale_prior=None,
                 model_class=ControlAffineExactGP):
        super().__init__()
        self.device = device or default_device()
        self.x_dim = x_dim
        self.u_dim = u_dim
        self.likelihood = IdentityLikelihood()
        self.model_class = model_class
        self.model = model_class(x_dim, u_dim, self.likelihood,
                                 gamma_length_scale_prior=gamma_length_scale_prior).to(device=self.device)
        self._cache = dict()
        self._jitter_source = None
        self._fit_noise_source = None
        self.fit_losses = []
        self._f_func_gp = GaussianProcess(self.f_func_mean, self.f_func_knl, (self.x_dim,), name="f")

    # ------------------------------------------------------------------------------------------ bookkeeping
    @property
    def ctrl_size(self):
        return self.u_dim

    @property
    def state_size(self):
        return self.x_dim

    def _params(self):
        """The model's parameters through cached (module, name) slots: the module tree is fixed after construction, and
        reading the slots sees re-assigned Parameter objects, unlike a cached list; `model.parameters()` walks the tree
        (~40 us here) and used to run several times per predict."""
        slots = getattr(self, '_param_slots', None)
        if slots is None or slots[0] is not self.model:
            found = []
            for mod in self.model.modules():
                for name in mod._parameters:
                    if mod._parameters[name] is not None:
                        found.append((mod, name))
            seen, uniq = set(), []
            for mod, name in found:
                prm = mod._parameters[name]
                if id(prm) not in seen:
                    seen.add(id(prm))
                    uniq.append((mod, name))
            self._param_slots = slots = (self.model, uniq)
        return [mod._parameters[name] for mod, name in slots[1]]

    @property
    def dtype(self):
        return self._params()[0].dtype

    def to(self, dtype=torch.float64):
        if dtype is torch.float64:
            self.double_()
        else:
            self.float_()

    def _cast(self, dt):
        self.model.to(dtype=dt)
        if self.model.train_inputs is not None:
            self.model.train_inputs = tuple(inp.to(dt) for inp in self.model.train_inputs)
            self.model.train_targets = self.model.train_targets.to(dt)
        # the factor cache is kept in float64 (the kernels compute in float64 whatever the model dtype)

    def double_(self):
        self._cast(torch.float64)
        assert self.dtype is torch.float64

    def float_(self):
        self._cast(torch.float32)
        assert self.dtype is torch.float32

    def set_jitter_source(self, vectors):
        """Explicit U(0,1) jitter vectors (consumed in order by every make_psd attempt) instead of torch.rand."""
        self._jitter_source = None if vectors is None else iter(vectors)

    def set_fit_noise_source(self, draws):
        """Explicit U(0,1) draws for the 1e-6 multiplicative target noise of every fit iteration (one array of
        N*n values per iteration, consumed in order) instead of torch.rand — parity tests replay the reference's draws."""
        self._fit_noise_source = None if draws is None else iter(draws)

    def _ensure_device_dtype(self, X):
        if isinstance(X, np.ndarray):
            X = torch.from_numpy(X)
        return X.to(device=self.device, dtype=self.dtype)

    def zero_grad(self):
        for p in self.model.parameters():
            if p.grad is not None:
                p.grad.detach_()
                p.grad.zero_()

    def clear_cache(self):
        self._cache = dict()

    # ------------------------------------------------------------------------------------------ hyper-parameters
    def _A_mat(self):
        return self.model.covar_module.task_covar_module.U.covar_matrix.evaluate()

    def _B_mat(self):
        return self.model.covar_module.task_covar_module.V.covar_matrix.evaluate()

    def _hyper64(self):
        """(lengthscale (n,), outputscale float, A, B, C (p,n)) detached, float64, on the device.  Cached until a
        parameter changes (in-place updates bump `_version`; re-assignment changes the object): the constrained values
        cost ~15 tiny launches and one device->host read (the outputscale), which used to be paid on every predict."""
        m = self.model
        key = tuple((id(p), p._version, p.dtype, p.device) for p in self._params())
        hc = getattr(self, '_hyper_cache', None)
        if hc is None or hc[0] != key:
            ls = m.input_covar.base_kernel.lengthscale.detach().reshape(-1).double().expand(self.x_dim).contiguous()
            s = float(m.input_covar.outputscale.detach())
            A = self._A_mat().detach().double().contiguous()
            B = self._B_mat().detach().double().contiguous()
            C = m.mean_module.constants().detach().double().contiguous()
            self._hyper_cache = hc = (key, (ls, s, A, B, C))
        return hc[1]

    def set_hyperparameters(self, lengthscale=None, outputscale=None, A=None, B=None, C=None):
        """Set the constrained hyper-parameters directly (lengthscale (n,), outputscale, A (n,n), B (p,p), C (p,n)).
        A / B are stored as a full-rank factor plus a small diagonal: chol(A - d I) chol(.)^T + d I with
        d = lambda_min(A) / 2.  Clears the factor cache."""
        from .gp_modules import inv_softplus
        m = self.model
        dev, dt = self.device, self.dtype
        t = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64) if not isinstance(v, torch.Tensor) else v,
                                      dtype=torch.float64)
        with torch.no_grad():
            if lengthscale is not None:
                m.input_covar.base_kernel.raw_lengthscale.copy_(inv_softplus(t(lengthscale)).reshape(1, -1).to(dev, dt))
            if outputscale is not None:
                m.input_covar.raw_outputscale.copy_(inv_softplus(t(outputscale)).reshape(()).to(dev, dt))
            for mat, ik in ((A, m.task_covar.U), (B, m.task_covar.V)):
                if mat is None:
                    continue
                M = t(mat)
                d = 0.5 * float(torch.linalg.eigvalsh(M).min())
                if not d > 0:
                    raise ValueError("set_hyperparameters: covariance matrix must be positive definite")
                Fm = torch.linalg.cholesky(M - d * torch.eye(M.shape[0], dtype=torch.float64))
                ik.covar_factor = nn.Parameter(Fm.to(dev, dt))
                ik.raw_var = nn.Parameter(inv_softplus(torch.full((M.shape[0],), d, dtype=torch.float64)).to(dev, dt))
            if C is not None:
                for bm, c in zip(m.mean_module.base_means, t(C).reshape(-1)):
                    bm.constant.fill_(float(c))
        self.clear_cache()
        return self

    def get_kernel_param(self, name):
        if name == 'A':
            return self._A_mat()
        elif name == 'B':
            return self._B_mat()
        elif name == 'scalefactor':
            return self.model.input_covar.outputscale
        elif name == 'lengthscale':
            return self.model.input_covar.base_kernel.lengthscale
        raise ValueError('Unknown param %s' % name)

    # ------------------------------------------------------------------------------------------ fit
    def fit(self, *args, max_cg_iterations=2000, **kwargs):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return self._fit_with_warnings(*args, **kwargs)

    def _fit_with_warnings(self, Xtrain_in, Utrain_in, XdotTrain_in, training_iter=50, lr=0.1, cuda_graph=None):
        """Adam + MultiStepLR on -(log marginal + lengthscale prior) / (N n) with a fresh 1e-6 multiplicative target noise
        per iteration (reference :274-335).  cuda_graph: None = capture value + gradients of one iteration in a CUDA graph
        when the problem is small enough to be launch-bound (Npad <= 1024, >= 8 iterations), True / False to force."""
        if Xtrain_in.shape[0] == 0:
            return self
        _need_cuda(self.device)
        model = self.model
        Xtrain, Utrain, XdotTrain = [self._ensure_device_dtype(X) for X in (Xtrain_in, Utrain_in, XdotTrain_in)]
        self.clear_cache()
        model.set_train_data(Xtrain, Utrain, XdotTrain)
        model.train()
        optimizer = torch.optim.Adam(model.parameters(), lr=lr)
        scheduler = torch.optim.lr_scheduler.MultiStepLR(
            optimizer, milestones=(torch.tensor([0.3, 0.6, 0.8, 0.90]) * training_iter).tolist())
        N, n = Xtrain.shape
        X64 = Xtrain.double().contiguous()
        UH64 = torch.cat([Xtrain.new_ones(N, 1), Utrain], dim=1).double().contiguous()
        prior = model.input_covar.base_kernel.lengthscale_prior
        params = list(model.parameters(recurse=True))
        self.fit_losses = []
        noise_buf = torch.zeros_like(XdotTrain)          # static input of the captured iteration

        def value_and_grads():
            """loss (0-d) and the fused finiteness flag of loss and gradients; gradients land in p.grad."""
            Y = (XdotTrain * (1 + 1e-6 * noise_buf)).double()
            ls = model.input_covar.base_kernel.lengthscale
            logp = mvgp_log_marginal(ls.double().reshape(-1).expand(n), model.input_covar.outputscale.double(),
                                     self._A_mat().double(), self._B_mat().double(),
                                     model.mean_module.constants().double(), X64, UH64, Y)
            if prior is not None:
                logp = logp + prior.log_prob(ls.double())
            loss = -logp / (N * n)
            loss.backward()
            # The reference asserts "no NaN" on every parameter, on the loss and on every gradient, each a device->host
            # read (~30 per iteration).  A NaN parameter makes the loss NaN, so one fused check of loss and gradients says
            # the same with ONE read per iteration.
            grads = [p.grad for p in params if p.grad is not None]
            ok = torch.isfinite(loss) & torch.isfinite(torch.stack(torch._foreach_norm(grads)).sum())
            return loss.detach(), ok

        graph = static = None
        want_graph = cuda_graph if cuda_graph is not None else (ops.padded(N) <= 1024 and training_iter >= 8)
        if want_graph and torch.device(self.device).type == 'cuda' and hasattr(torch.cuda, 'CUDAGraph'):
            graph, static = self._capture_fit_iteration(value_and_grads, optimizer)
        for i in range(training_iter):
            # fresh multiplicative target noise every iteration (reference :318-321), drawn on the CPU generator
            if self._fit_noise_source is not None:
                noise = torch.as_tensor(next(self._fit_noise_source)).reshape(XdotTrain.shape).to(dtype=XdotTrain.dtype)
            else:
                noise = torch.rand(XdotTrain.shape, dtype=XdotTrain.dtype)
            noise_buf.copy_(noise)
            replayed = False
            if graph is not None:
                graph.replay()
                loss, status, _ = static
                bad_pivot, not_finite = status.tolist()   # the one device->host read of the iteration
                replayed = bad_pivot == 0
            if not replayed:                              # eager iteration (with the psd-safe jitter escalation)
                optimizer.zero_grad(set_to_none=False)    # keep the gradient tensors: the graph writes to these
                loss, ok = value_and_grads()
                not_finite = not bool(ok)
            assert not not_finite, "NaN / inf in the loss or a gradient of the log marginal likelihood"
            self.fit_losses.append(loss.clone())
            if LOG.isEnabledFor(logging.DEBUG):
                LOG.debug('Iter %d/%d - Loss: %.3f' % (i + 1, training_iter, loss.item()))
            optimizer.step()
            scheduler.step()
        return self

    def _capture_fit_iteration(self, value_and_grads, optimizer):
        """One iteration's value + gradients as a CUDA graph (the small-N regime is launch- and Python-bound: ~150 tiny
        launches per iteration).  Returns (graph, (loss, status, info)) with static tensors — status = [Cholesky info,
        not-finite flag], info the status word the factor kernels write — or (None, None) when the capture is not possible:
        the caller then iterates eagerly.  The optimiser step stays outside the graph."""
        from .mll import capturable
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side), capturable(info):
                optimizer.zero_grad(set_to_none=True)     # warm-up: lazy initialisations, scratch growth, cuBLAS handles
                value_and_grads()
                optimizer.zero_grad(set_to_none=True)
                side.synchronize()
                # capture_begin / capture_end directly: torch.cuda.graph() would also run the garbage collector and empty
                # the caching allocator, which costs more than the iterations a short fit saves
                graph.capture_begin()
                try:
                    loss, ok = value_and_grads()
                    status = torch.stack([info[0].to(torch.int64), (~ok).to(torch.int64)])
                finally:
                    graph.capture_end()
            torch.cuda.current_stream().wait_stream(side)
            # `info` is written by the graph on every replay (a memset + the factor kernels): it must outlive this call
            return graph, (loss, status, info)
        except Exception as e:                            # e.g. an allocation inside the capture: stay eager
            LOG.warning("fit: CUDA-graph capture of the iteration failed (%s); iterating eagerly" % (e,))
            optimizer.zero_grad(set_to_none=True)
            torch.cuda.synchronize()
            return None, None

    # ------------------------------------------------------------------------------------------ factor
    def _train_data(self):
        MXUHtrain = self.model.train_inputs[0]
        _, Xtrain, UHtrain = self.model.decoder.decode(MXUHtrain)
        N = Xtrain.size(0)
        return Xtrain, UHtrain, self.model.train_targets.reshape(N, -1)

    def _train_data64(self):
        """(X, UH, targets) of the train set as contiguous float64 tensors, kept until the train set changes (a new tensor
        or an in-place write): derived from the DATA, not from the factorisation, so `clear_cache()` does not drop them —
        the reference's speed test clears the cache inside its timed statement (pendulum.py:1367-1372) and these three
        casts were ~6 launches of every call."""
        inp, tgt = self.model.train_inputs[0], self.model.train_targets
        key = (id(inp), inp._version, id(tgt), tgt._version)
        tc = getattr(self, '_train64_cache', None)
        if tc is None or tc[0] != key:
            Xtrain, UHtrain, targets = self._train_data()
            tc = (key, (Xtrain.double().contiguous(), UHtrain.double().contiguous(), targets.double().contiguous()),
                  (inp, tgt))        # the tensors are held so that their ids cannot be recycled
            self._train64_cache = tc
        return tc[1]

    def _perturbed_cholesky_compute(self, k, B, Xtrain, UHtrain, cholesky_tries=10, cholesky_perturb_init=1e-5,
                                    cholesky_perturb_scale=10):
        """Kb = k(X,X) o (UH B UH^T); L = chol(Kb + 1e-5 * 10^t * diag(U(0,1)))  (reference :366-377, :899-921).
        `k` is accepted for signature compatibility; the data kernel's hyper-parameters are read from the model."""
        _need_cuda(Xtrain.device)
        ls, s, _, _, _ = self._hyper64()
        tr = self._train_data()
        if Xtrain is tr[0] or (Xtrain.data_ptr() == tr[0].data_ptr() and Xtrain.shape == tr[0].shape
                               and UHtrain.data_ptr() == tr[1].data_ptr()):
            X64, UH64, _ = self._train_data64()
        else:
            X64, UH64 = Xtrain.double().contiguous(), UHtrain.double().contiguous()
        B64 = B.detach().double().contiguous()
        N = X64.shape[0]
        factor = cholesky_perturb_init
        for ntry in range(cholesky_tries):
            eps = next(self._jitter_source) if self._jitter_source is not None else _draw_jitter(N, self.dtype)
            eps = eps.to(device=X64.device, dtype=torch.float64).contiguous()
            Kb = ops.gram_train_lower(X64, UH64, B64, ls, s)
            try:
                # the inverse is queued BEFORE the factor's status is read back (the read synchronises): a failed factor
                # is all-NaN, its inverse is harmless, and the device does not idle while the host comes back
                info = torch.zeros(1, dtype=torch.int32, device=X64.device)
                L, dinv = ops.potrf_(Kb, N, eps, factor, check_pd=False, info_out=info)
                Linv = ops.trtri(L, dinv)
                ops.check_info(info)
                break
            except RuntimeError as e:
                if ntry == cholesky_tries - 1:
                    raise
                LOG.warning("Cholesky failed with perturb={} on error {}".format(factor, str(e)))
                factor = factor * cholesky_perturb_scale
        self._cache['_Lpad'] = L
        self._cache['_Linv'] = Linv
        self._cache['_jitter'] = (eps, factor)          # what was added to the diagonal: the alpha refinement needs it
        return L[:N, :N]

    def _perturbed_cholesky(self, k, B, Xtrain, UHtrain, cache_key="perturbed_cholesky"):
        if cache_key not in self._cache:
            self._cache[cache_key] = self._perturbed_cholesky_compute(k, B, Xtrain, UHtrain)
        return self._cache[cache_key]

    def _factor_state(self):
        """Cached (Linv (Npad,Npad), alpha (Npad,n), G (Npad,p), Y (N,n)) for the current train data."""
        Xtrain, UHtrain, targets = self._train_data()
        ls, s, A, B, C = self._hyper64()
        self._perturbed_cholesky(None, B, Xtrain, UHtrain)
        if '_alpha' not in self._cache:
            Linv = self._cache['_Linv']
            Npad, N = Linv.shape[0], Xtrain.shape[0]
            X64, UH64, T64 = self._train_data64()
            Y = T64 - UH64 @ C                                        # Y = Xdot - UH C  (:525-532)
            Ypad = torch.zeros(Npad, Y.shape[1], dtype=torch.float64, device=Y.device)
            Ypad[:N] = Y
            # Kb^-1 Y (:545, cholesky_solve there): explicit-inverse product + three refinement steps whose residual is taken
            # against the factorised matrix itself in compensated arithmetic (bcbf_alpha_refine)
            eps, factor = self._cache['_jitter']
            self._cache['_alpha'] = ops.alpha_refine(X64, UH64, B, ls, s, Linv, Ypad, eps, factor, iters=3).contiguous()
            G = torch.zeros(Npad, B.shape[0], dtype=torch.float64, device=Y.device)
            G[:N] = UH64 @ B
            self._cache['_G'] = G
        return self._cache['_Linv'], self._cache['_alpha'], self._cache['_G']

    # ------------------------------------------------------------------------------------------ prediction
    def _uh(self, X, U_in, fill):
        if U_in is None:
            UH = X.new_zeros(X.shape[0], self.model.matshape[0])
            UH[:, 0] = 1
            return UH
        U = self._ensure_device_dtype(U_in)
        return torch.cat((U.new_full((U.shape[0], 1), fill), U), dim=-1)

    def _kb(self, X64, UH64, Xq, UHq, ls, s, B, Npad, diff):
        """k(Xtrain, Xq) o (UHtrain B UHq^T), zero-padded to Npad rows."""
        if diff:
            K = autograd_ops.rbf_kernel(X64, Xq, ls, torch.as_tensor(s, dtype=torch.float64, device=Xq.device))
            kb = K * ((UH64 @ B) @ UHq.transpose(0, 1))
            return torch.nn.functional.pad(kb, (0, 0, 0, Npad - kb.shape[0]))
        return ops.gram_ca(X64, Xq.contiguous(), ls, s, UH64, UHq.contiguous(), B, rows_pad=Npad)

    def _kss(self, Xq, UHq, Xp, UHp, ls, s, B, diff):
        if diff:
            K = autograd_ops.rbf_kernel(Xq, Xp, ls, torch.as_tensor(s, dtype=torch.float64, device=Xq.device))
            return K * (UHq @ B @ UHp.transpose(0, 1))
        return ops.gram_ca(Xq.contiguous(), Xp.contiguous(), ls, s, UHq.contiguous(), UHp.contiguous(), B)

    def custom_predict(self, Xtest_in, Utest_in=None, UHfill=1, Xtestp_in=None, Utestp_in=None, UHfillp=1,
                       compute_cov=True, grad_gp=False, grad_check=False, scalar_var_only=False):
        """Posterior of F(x)[UHfill;u] with u folded in before the solve (reference :390-613, R&W Alg. 2.1):
        returns (mean (b,n), cov (1, b*n, b'*n) = scalar_var (x) A)  [or scalar_var (b,b') if scalar_var_only]."""
        if grad_gp:
            return self._custom_predict_grad(Xtest_in, Utest_in, UHfill, Xtestp_in, Utestp_in, UHfillp, compute_cov,
                                             scalar_var_only)
        _need_cuda(self.device)
        Xtest = self._ensure_device_dtype(Xtest_in)
        Xtestp = self._ensure_device_dtype(Xtestp_in) if Xtestp_in is not None else Xtest
        UHtest = self._uh(Xtest, Utest_in, UHfill)
        UHtestp = UHtest if Utestp_in is None else self._uh(Xtestp, Utestp_in, UHfillp)
        out_dt = self.dtype
        ls, s, A, B, C = self._hyper64()
        Xq, Xp, UHq, UHp = Xtest.double(), Xtestp.double(), UHtest.double(), UHtestp.double()
        diff = any(t.requires_grad for t in (Xq, Xp, UHq, UHp))
        fu_mean_test = UHq @ C                                              # (b, n): M(x)^T [1;u]
        if self.model.train_inputs is None:
            scalar_var = self._kss(Xq, UHq, Xp, UHp, ls, s, B, diff)
            return fu_mean_test.to(out_dt), torch_kron(scalar_var.unsqueeze(0), A.unsqueeze(0)).to(out_dt)
        Xtrain, UHtrain, _ = self._train_data()
        X64, UH64 = Xtrain.double().contiguous(), UHtrain.double().contiguous()
        Linv, alpha, _ = self._factor_state()
        Npad = Linv.shape[0]
        kb_star = self._kb(X64, UH64, Xq, UHq, ls, s, B, Npad, diff)       # (Npad, b)  (:536)
        if diff:
            mean = fu_mean_test + autograd_ops.mm_tn(kb_star, alpha)       # (:547)
        else:
            mean = fu_mean_test + ops.gemm(kb_star, alpha, transa=True)
        if not compute_cov:
            return mean.to(out_dt), (0 * A).to(out_dt)                      # (:612)
        kb_star_p = self._kb(X64, UH64, Xp, UHp, ls, s, B, Npad, diff) if Xtestp_in is not None else kb_star
        kb_ss = self._kss(Xq, UHq, Xp, UHp, ls, s, B, diff)                # (b, b')
        if diff:
            v = autograd_ops.linv_mm(Linv, kb_star)                        # L \ kb*  (:565)
            vp = autograd_ops.linv_mm(Linv, kb_star_p) if Xtestp_in is not None else v
            scalar_var = kb_ss - autograd_ops.mm_tn(v, vp)                 # (:586)
        else:
            v = ops.trmm_lower(Linv, kb_star.contiguous())
            vp = ops.trmm_lower(Linv, kb_star_p.contiguous()) if Xtestp_in is not None else v
            scalar_var = ops.gemm(v, vp, transa=True, alpha=-1.0, beta=1.0, C=kb_ss)
        if scalar_var_only:
            return mean.to(out_dt), scalar_var.to(out_dt)
        return mean.to(out_dt), torch_kron(scalar_var.unsqueeze(0), A.unsqueeze(0)).to(out_dt)   # (:602)

    def _custom_predict_grad(self, Xtest_in, Utest_in, UHfill, Xtestp_in, Utestp_in, UHfillp, compute_cov,
                             scalar_var_only):
        """custom_predict(grad_gp=True): posterior of the GRADIENT process d/dx [F(x)[UHfill;u]] (reference :447-477).

        mean (b, n*n): entry [t, r*n + d] = d mean_r(x_t; u_t) / d x_{t,d}  (the layout of the reference's
        `fu_mean_test` reshape, :486-493); covariance: the scalar part is d^2/dx dx' [kb**(x,x') - v(x)^T v(x')],
        a (b*n, b'*n) matrix with rows (t, d) and columns (t', d'), Kronecker-expanded with A like the value process.
        Closed-form derivative kernel blocks from bcbf_rbf_blocks replace the reference's autograd closures
        grad_ksx / grad_kxs / Hessian_kxx.  Deviation, on purpose: the reference's grad_kxs calls
        autograd.grad(list(k(Xtrain, x*)), x*), which SUMS the kernel gradient over the training points before it is
        weighted by (uh_i B uh*) and alpha_i; that expression is not the gradient of anything, its one caller is
        commented out (tests/test_control_affine_regression.py:184,191), and for n > 1 it stops on a shape error.
        This method returns what that commented-out test compares against: the gradient of `fu_func_mean`."""
        _need_cuda(self.device)
        Xtest = self._ensure_device_dtype(Xtest_in)
        Xtestp = self._ensure_device_dtype(Xtestp_in) if Xtestp_in is not None else Xtest
        UHtest = self._uh(Xtest, Utest_in, UHfill)
        UHtestp = UHtest if Utestp_in is None else self._uh(Xtestp, Utestp_in, UHfillp)
        out_dt = self.dtype
        ls, s, A, B, C = self._hyper64()
        Xq, Xp, UHq, UHp = Xtest.double().contiguous(), Xtestp.double().contiguous(), UHtest.double(), UHtestp.double()
        b, n = Xq.shape
        bp_ = Xp.shape[0]

        def hess_prior():                       # d^2 k(x_t, x'_t') / dx dx'^T * (uh_t B uh'_t'): rows (t,d), cols (t',d')
            _, _, d2K = ops.rbf_blocks(Xq, Xp, ls, s, hess=True)                    # (b, b', n, n)
            w = (UHq @ B @ UHp.transpose(0, 1)).unsqueeze(-1).unsqueeze(-1)
            return (d2K * w).permute(0, 2, 1, 3).reshape(b * n, bp_ * n)

        mean = Xq.new_zeros(b, n * n)           # the constant prior mean has zero gradient (:452-461)
        if self.model.train_inputs is None:
            scalar_var = hess_prior()
            cov = scalar_var if scalar_var_only else torch_kron(scalar_var.unsqueeze(0), A.unsqueeze(0))
            return mean.to(out_dt), cov.to(out_dt)
        Xtrain, UHtrain, _ = self._train_data()
        X64 = Xtrain.double().contiguous()
        N = X64.shape[0]
        Linv, alpha, G = self._factor_state()
        Npad = Linv.shape[0]

        def dkb(Xs, UHs):                       # d kb*(x_t)[i] / d x_{t,d}  as an (Npad, b*n) matrix, columns (t, d)
            _, dK, _ = ops.rbf_blocks(Xs, X64, ls, s, grad=True)                    # (b, N, n), derivative w.r.t. x_t
            w = (G[:N] @ UHs.transpose(0, 1))                                       # (N, b): uh_i^T B uh_t
            M = (dK.permute(1, 0, 2) * w.unsqueeze(-1)).reshape(N, -1)
            return torch.nn.functional.pad(M, (0, 0, 0, Npad - N)).contiguous()

        dkb_q = dkb(Xq, UHq)
        mean = ops.gemm(dkb_q, alpha, transa=True).reshape(b, n, -1).transpose(1, 2).reshape(b, -1)   # [t, r*n + d]
        if not compute_cov:
            return mean.to(out_dt), (0 * A).to(out_dt)
        dkb_p = dkb(Xp, UHp) if (Xtestp_in is not None or Utestp_in is not None) else dkb_q
        V = ops.trmm_lower(Linv, dkb_q)
        Vp = ops.trmm_lower(Linv, dkb_p) if dkb_p is not dkb_q else V
        scalar_var = ops.gemm(V, Vp, transa=True, alpha=-1.0, beta=1.0, C=hess_prior())
        if scalar_var_only:
            return mean.to(out_dt), scalar_var.to(out_dt)
        return mean.to(out_dt), torch_kron(scalar_var.unsqueeze(0), A.unsqueeze(0)).to(out_dt)

    # ---- matrix form (class Exact in the reference; kept on the base class so that predict() can use it) ----------
    def _custom_predict_matrix(self, Xtest_in, Xtestp_in=None, compute_cov=True, _out_jitter=True, _pending=None,
                               _psd_start_try=0):
        """M_k (b,n,p), A (n,n), B_k (b,b',p,p)  (reference :983-1096).  _pending: a list; when given, the status read of
        the output make_psd is deferred and its check appended for the caller to run last (custom_predict_fullmat)."""
        _need_cuda(self.device)
        Xtest = self._ensure_device_dtype(Xtest_in)
        Xtestp = self._ensure_device_dtype(Xtestp_in) if Xtestp_in is not None else Xtest
        out_dt = self.dtype
        ls, s, A, B, C = self._hyper64()
        p, n = self.model.matshape
        Xq, Xp = Xtest.double(), Xtestp.double()
        b, bp_ = Xq.shape[0], Xp.shape[0]
        diff = Xq.requires_grad or Xp.requires_grad
        if diff:
            s_t = torch.as_tensor(s, dtype=torch.float64, device=Xq.device)
        kfun = (lambda a, c: autograd_ops.rbf_kernel(a, c, ls, s_t)) if diff else \
            (lambda a, c: ops.gram_ca(a.contiguous(), c.contiguous(), ls, s))
        M0 = C.t().unsqueeze(0).expand(b, n, p)                             # (:1022-1023)
        if self.model.train_inputs is None:
            return M0.to(out_dt), A.to(out_dt), (B * kfun(Xq, Xp).unsqueeze(-1).unsqueeze(-1)).to(out_dt)
        X64, UH64, _ = self._train_data64()
        N = X64.shape[0]
        Linv, alpha, G = self._factor_state()
        Npad = Linv.shape[0]
        # frakB[i, (t, q)] = k(X_i, x_t) G[i, q]   (Npad, b*p)   (:1051)
        if diff:
            K = kfun(X64, Xq)                                               # (N, b)
            frakB = (K.unsqueeze(-1) * G[:N].unsqueeze(1)).reshape(N, b * p)
            frakB = torch.nn.functional.pad(frakB, (0, 0, 0, Npad - N))
            mean_k = M0 + autograd_ops.mm_tn(alpha, frakB).reshape(n, b, p).permute(1, 0, 2)
        else:
            Xr, E = HetergeneousMatrixVariateKernel._onehot_cols(Xq, p)
            frakB = ops.gram_ca(X64, Xr, ls, s, UH64, E, B, rows_pad=Npad)
            mean_k = M0 + ops.gemm(alpha, frakB.contiguous(), transa=True).reshape(n, b, p).permute(1, 0, 2)  # (:1055)
        if not compute_cov:
            return mean_k.to(out_dt), A.to(out_dt), Xtest.new_zeros(b, bp_, p, p)
        KB = torch_kron(kfun(Xq, Xp), B, batch_dims=0)                      # (:1062-1063)
        # the reference subtracts the Xtest term on both sides even when Xtestp is given (:1079-1088; needs b == b')
        if diff:
            V = autograd_ops.linv_mm(Linv, frakB)
            BkXX = KB - autograd_ops.mm_tn(V, V)
        else:
            V = ops.trmm_lower(Linv, frakB.contiguous())
            BkXX = ops.gemm(V, V, transa=True, alpha=-1.0, beta=1.0, C=KB)
        if _out_jitter and _pending is not None and not diff:
            BkXX, _, chk = self._make_psd_output(BkXX, start_try=_psd_start_try, defer=True)
            _pending.append(chk)
        elif _out_jitter:
            BkXX, _ = self._make_psd_output(BkXX, start_try=_psd_start_try)  # (:1089)
        BkXX = BkXX.reshape(b, p, bp_, p).transpose(1, 2)                   # (:1091)
        return mean_k.to(out_dt), A.to(out_dt), BkXX.to(out_dt)

    def _make_psd_output(self, M, start_try=0, defer=False):
        """The reference's second make_psd (:1089): random jitter ADDED to the returned covariance, retried x10 (factor
        1e-5 * 10^t) until the blocked Cholesky accepts it.  defer=True: the first attempt's status is not read back here;
        returns (Mp, L, pending) with `pending()` -> True when that attempt succeeded — the caller queues the rest of its
        work first and reads the status last, so the device does not idle behind the read (a failed attempt is redone by
        the caller with start_try + 1)."""
        n = M.shape[0]
        factor = 1e-5 * 10 ** start_try
        for ntry in range(start_try, 10):
            eps = next(self._jitter_source) if self._jitter_source is not None else _draw_jitter(n, self.dtype)
            eps = eps.to(device=M.device, dtype=torch.float64)
            Mp = M + factor * torch.diag(eps)
            buf = torch.eye(ops.padded(n), dtype=torch.float64, device=M.device)
            buf[:n, :n] = Mp.detach()
            if defer:
                info = torch.zeros(1, dtype=torch.int32, device=M.device)
                L, _ = ops.potrf_(buf, n, None, 0.0, check_pd=False, info_out=info)

                def pending(factor=factor, info=info):
                    try:
                        ops.check_info(info)
                        return True
                    except RuntimeError as e:
                        LOG.warning("Cholesky failed with perturb={} on error {}".format(factor, str(e)))
                        return False
                return Mp, L[:n, :n], pending
            try:
                L, _ = ops.potrf_(buf, n, None, 0.0)
                return Mp, L[:n, :n]
            except RuntimeError as e:
                if ntry == 9:
                    raise
                LOG.warning("Cholesky failed with perturb={} on error {}".format(factor, str(e)))
                factor *= 10
        raise AssertionError("unreachable")

    def custom_predict_blocks(self, Xtest_in, Utest_in=None):
        """Per-query posterior blocks for large batches (extension; the reference's (b,b,p,p) form is O(b^2)):
        M_k (b,n,p), B_k (b,p,p) [no output jitter] and, when Utest is given, mean (b,n), svar (b,) = u^T B_k u.
        Fused path: cross Gram -> posterior kernel.  `self.covariance_kernel` picks the N^2 p contraction: 'dmma' = FP64
        tensor pipe (bcbf_posterior_blocks), 'int8' = tcgen05 int8 tensor cores with error-free digit splitting
        (bcbf_posterior_blocks_i8, same result to FP64 rounding, ~3x the throughput at large N), 'auto' (default) =
        int8 once the problem is large enough to fill the machine (N >= 1024 and >= 512 queries)."""
        _need_cuda(self.device)
        Xq = self._ensure_device_dtype(Xtest_in).double().contiguous()
        ls, s, A, B, C = self._hyper64()
        p, n = self.model.matshape
        Xtrain, _, _ = self._train_data()
        Linv, alpha, G = self._factor_state()
        if '_W' not in self._cache:
            self._cache['_W'] = (alpha.unsqueeze(-1) * G.unsqueeze(1)).reshape(G.shape[0], n * p).contiguous()
        Q = Xq.shape[0]
        Ks = ops.cross_gram(Xtrain.double().contiguous(), Xq, ls, s, Npad=Linv.shape[0])
        kern = getattr(self, 'covariance_kernel', 'auto')
        Npad = Linv.shape[0]
        if kern == 'auto':
            kern = 'int8' if (Npad >= 1024 and Q >= 512 and Npad <= ops.oz_max_npad()) else 'dmma'
        if kern == 'int8':
            if '_oz' not in self._cache:
                self._cache['_oz'] = ops.oz_split_factor(Linv)
            digits, rowscale = self._cache['_oz']
            Mk, Bk = ops.posterior_blocks_i8(digits, rowscale, Ks, G, self._cache['_W'], B, C.t().contiguous(), s, n, p, Q)
        else:
            Mk, Bk = ops.posterior_blocks(Linv, Ks, G, self._cache['_W'], B, C.t().contiguous(), s, n, p, Q)
        if Utest_in is None:
            return Mk, Bk
        UHq = self._uh(Xq, Utest_in, 1).double().contiguous()
        mean, svar = ops.contract_u(Mk, Bk, UHq)
        return Mk, Bk, mean, svar

    def custom_predict_fullmat(self, Xtest_in, Xtestp_in=None):
        """vec F(x) in (b,p,n) order and its (bpn, bpn) covariance (reference :963-980)."""
        Xtest = self._ensure_device_dtype(Xtest_in)
        b, p = Xtest.shape[0], 1 + self.u_dim
        for start_try in range(10):
            # every launch of the call is queued before the two status reads (output make_psd, NaN check): each read
            # synchronises, and the device would otherwise idle while the host queues what follows
            pending = []
            meanFX, A, BkXX = self._custom_predict_matrix(Xtest_in, Xtestp_in, compute_cov=True, _pending=pending,
                                                          _psd_start_try=start_try)
            assert meanFX.shape == (b, self.x_dim, p)
            nan_mean = torch.isnan(meanFX).any()
            mean_out = meanFX.transpose(-2, -1).reshape(-1)
            var_FX = torch_kron(BkXX.transpose(2, 1).reshape(b * p, b * p), A, batch_dims=0)
            if all(chk() for chk in pending):
                break
            if start_try == 9:
                raise RuntimeError("make_psd: the output covariance is not positive definite after 10 perturbations")
        assert not bool(nan_mean)
        return mean_out, var_FX

    def predict(self, Xtest_in, return_cov=True):
        """mean F(x)^T (b,p,n) and covariance (bpn,bpn) on the INPUT's device / dtype (reference :343-364; there the
        numbers come from gpytorch's eval-mode ExactGP — parity unpinned; here the closed form)."""
        Xtest = self._ensure_device_dtype(Xtest_in)
        if isinstance(Xtest_in, np.ndarray):
            Xtest_in = torch.from_numpy(Xtest_in)
        meanFX, A, BkXX = self._custom_predict_matrix(Xtest, None, compute_cov=return_cov, _out_jitter=False)
        b, p = Xtest.shape[0], 1 + self.u_dim
        mean = meanFX.transpose(-2, -1).to(device=Xtest_in.device, dtype=Xtest_in.dtype)
        if not return_cov:
            return mean
        cov = torch_kron(BkXX.transpose(2, 1).reshape(b * p, b * p), A, batch_dims=0)
        return mean, cov.to(device=Xtest_in.device, dtype=Xtest_in.dtype)

    # ------------------------------------------------------------------------------------------ closures
    @staticmethod
    def _b(t):
        return t.unsqueeze(0) if t.ndim == 1 else t

    def f_func(self, Xtest_in, return_cov=False):
        Xtest = self._b(Xtest_in)
        Utest = Xtest.new_zeros((Xtest.shape[0], self.u_dim))
        mean_fx, cov_fx = self.custom_predict(Xtest, Utest)
        if return_cov:
            if Xtest_in.ndim == 1:
                cov_fx = cov_fx.squeeze(0)
            cov_fx = cov_fx.to(dtype=Xtest_in.dtype, device=Xtest_in.device)
        if Xtest_in.ndim == 1:
            mean_fx = mean_fx.squeeze(0)
        mean_fx = mean_fx.to(dtype=Xtest_in.dtype, device=Xtest_in.device)
        return (mean_fx, cov_fx) if return_cov else mean_fx

    def f_func_mean(self, Xtest_in):
        mean_f, _ = self.custom_predict(self._b(Xtest_in), compute_cov=False)
        if Xtest_in.ndim == 1:
            mean_f = mean_f.squeeze(0)
        return mean_f.to(dtype=Xtest_in.dtype, device=Xtest_in.device)

    def f_func_knl(self, Xtest_in, Xtestp_in, grad_check=False):
        _, var_f = self.custom_predict(self._b(Xtest_in), Xtestp_in=self._b(Xtestp_in), compute_cov=True)
        if Xtest_in.ndim == 1:
            var_f = var_f.squeeze(0)
        return var_f.to(dtype=Xtest_in.dtype, device=Xtest_in.device)

    def f_func_gp(self):
        return self._f_func_gp

    def fu_func_mean(self, Utest_in, Xtest_in):
        mean_f, _ = self.custom_predict(self._b(Xtest_in), self._b(Utest_in), compute_cov=False)
        if Xtest_in.ndim == 1:
            mean_f = mean_f.squeeze(0)
        return mean_f.to(dtype=Xtest_in.dtype, device=Xtest_in.device)

    def _grad_fu_func_mean(self, Xtest_in, Utest_in=None):
        """d/dx of the posterior mean of F(x)[1;u] (reference :759-771): (n*n,) for a single state, (b, n*n) batched,
        entry r*n + d = d mean_r / d x_d."""
        Utest = self._b(Utest_in) if Utest_in is not None else None
        mean_f, _ = self.custom_predict(self._b(Xtest_in), Utest, compute_cov=False, grad_gp=True)
        if Xtest_in.ndim == 1:
            mean_f = mean_f.squeeze(0)
        return mean_f.to(dtype=Xtest_in.dtype, device=Xtest_in.device)

    def fu_func_knl(self, Utest_in, Xtest_in, Xtestp_in):
        _, var_f = self.custom_predict(self._b(Xtest_in), self._b(Utest_in), Xtestp_in=self._b(Xtestp_in),
                                       compute_cov=True)
        if Xtest_in.ndim == 1:
            var_f = var_f.squeeze(0)
        return var_f.to(dtype=Xtest_in.dtype, device=Xtest_in.device)

    def fu_func_gp(self, Utest_in):
        gp = GaussianProcess(mean=partial(self.fu_func_mean, Utest_in), knl=partial(self.fu_func_knl, Utest_in),
                             shape=(self.x_dim,), name="F(.)u")
        gp.register_covar(self._f_func_gp, partial(self.covar_fu_f, Utest_in))
        return gp

    def covar_fu_f(self, Utest_in, Xtest_in, Xtestp_in):
        Utest = self._b(Utest_in)
        _, var_f = self.custom_predict(self._b(Xtest_in), Utest, Xtestp_in=self._b(Xtestp_in),
                                       Utestp_in=torch.zeros_like(Utest), compute_cov=True)
        if Xtest_in.ndim == 1:
            var_f = var_f.squeeze(0)
        return var_f.to(dtype=Xtest_in.dtype, device=Xtest_in.device)

    def g_func(self, Xtest_in, return_cov=False):
        assert not return_cov, "Don't know what matrix covariance looks like"
        mean_Fx = self.predict(self._b(Xtest_in), return_cov=False)
        mean_gx = mean_Fx[:, 1:, :]
        if Xtest_in.ndim == 1:
            mean_gx = mean_gx.squeeze(0)
        return mean_gx.to(dtype=Xtest_in.dtype, device=Xtest_in.device).transpose(-2, -1)

    def _gu_func(self, Xtest_in, Utest_in=None, return_cov=False, Xtestp_in=None):
        Xtest = self._b(Xtest_in)
        Utest = self._b(Utest_in) if Utest_in is not None else Xtest_in.new_ones(Xtest.shape[0], self.u_dim)
        mean_gu, var_gu = self.custom_predict(Xtest, Utest, UHfill=0, Xtestp_in=Xtestp_in, compute_cov=True)
        if Xtest_in.ndim == 1 and (Utest_in is None or Utest_in.ndim == 1):
            mean_gu = mean_gu.squeeze(0)
            var_gu = var_gu.squeeze(0)
        return (mean_gu, var_gu) if return_cov else mean_gu

    def g_func_mean(self, Xtest_in):
        return self._gu_func(Xtest_in, return_cov=False)

    def _predict_flatten(self, Xtest_in, Utest_in):
        """f(x, u) = f(x) + g(x) u for train-style rows (mask 1): mean (b, n) and covariance reshaped (b, n, n, b) as the
        reference does with gpytorch's eval-mode output (:645-683, the raw reshape of the (b n, b n) matrix included).
        The numbers are the closed-form posterior (custom_predict) — eval-mode gpytorch is parity-unpinned."""
        if isinstance(Xtest_in, np.ndarray):
            Xtest_in = torch.from_numpy(Xtest_in)
        if isinstance(Utest_in, np.ndarray):
            Utest_in = torch.from_numpy(Utest_in)
        if self.model is None or self.likelihood is None:
            raise RuntimeError("Call fit() with training data before calling predict")
        Xtest = self._ensure_device_dtype(Xtest_in)
        mean, cov = ControlAffineRegressor.custom_predict(self, Xtest, self._ensure_device_dtype(Utest_in))
        b, n = Xtest.shape[0], self.x_dim
        cov = cov.reshape(b * n, b * n).reshape(b, n, n, b)
        return (mean.to(device=Xtest_in.device, dtype=Xtest_in.dtype), cov.to(device=Xtest_in.device, dtype=Xtest_in.dtype))

    def _cbf_func(self, Xtest, grad_htest, return_cov=False):
        """grad_h F(x) and, on request, its variance grad_h^T cov(F) grad_h (reference :853-860; there the
        return_cov=False branch unpacks a single tensor and fails — here it returns (mean, None))."""
        if return_cov:
            mean_Fx, cov_Fx = self.predict(Xtest, return_cov=True)
            cov_hFT = grad_htest.T @ cov_Fx @ grad_htest
        else:
            mean_Fx, cov_hFT = self.predict(Xtest, return_cov=False), None
        return grad_htest @ mean_Fx, cov_hFT

    # ------------------------------------------------------------------------------------------ persistence
    def state_dict(self):
        return dict(model=self.model.state_dict(), likelihood=dict())

    def load_state_dict(self, state_dict):
        self.model.load_state_dict(state_dict['model'])
        self.clear_cache()      # the cached factor / alpha / W belong to the previous parameters and train data

    def save(self, path='/tmp/saved.pickle'):
        torch.save(self.state_dict(), path)

    def load(self, path='/tmp/saved.pickle'):
        self.load_state_dict(torch.load(path, weights_only=False))


ControlAffineRegressorRankOne = partial(
    ControlAffineRegressor,
    model_class=partial(ControlAffineExactGP, rank=1, gamma_length_scale_prior=(1e-3, 1e-3)))


class ControlAffineRegressorExact(ControlAffineRegressor):
    """Matrix-variate form: posterior of F(x) first, then the [1;u] contraction (reference :930-1096)."""

    def custom_predict(self, Xtest_in, Utest_in=None, UHfill=1, Xtestp_in=None, Utestp_in=None, UHfillp=1,
                       compute_cov=True):
        Xtest = self._ensure_device_dtype(Xtest_in)
        Xtestp = self._ensure_device_dtype(Xtestp_in) if Xtestp_in is not None else Xtest
        meanFX, A, BkXX = self._custom_predict_matrix(Xtest_in, Xtestp_in, compute_cov=compute_cov)
        UHtest = self._uh(Xtest, Utest_in, UHfill)
        UHtestp = UHtest if Utestp_in is None else self._uh(Xtestp, Utestp_in, UHfillp)
        meanFXU = meanFX.bmm(UHtest.unsqueeze(-1)).squeeze(-1)
        if compute_cov:
            l = UHtest.unsqueeze(-1).unsqueeze(1)       # (b, 1, p, 1)
            r = UHtestp.unsqueeze(-1).unsqueeze(0)      # (1, b', p, 1)
            varFXU = torch.matmul(torch.matmul(l.transpose(-2, -1), BkXX), r) * A
        else:
            varFXU = Xtest.new_zeros(Xtest.shape[0], Xtestp.shape[0], *A.shape)
        return meanFXU, varFXU


ControlAffineRegressorExactRankOne = partial(
    ControlAffineRegressorExact,
    model_class=partial(ControlAffineExactGP, rank=1))

# rank-0 ("diagonal") variant of the task covariances (reference :1334-1336)
ControlAffineRegMatrixDiag = partial(
    ControlAffineRegressorExact,
    model_class=partial(ControlAffineExactGP, rank=0))


# =====================================================================================================================
# CoGP comparator: vec F(x) ~ GP(vec M(x), Sigma k(x,x')) with one (p n) x (p n) coregionalisation matrix
# instead of the Kronecker pair (A, B).  It exists in the reference as the baseline the MVGP is compared against
# (control_affine_model.py:1106-1331; speed test pendulum.py:1316-1319).  The (N n) x (N n) factorisation and every
# N-sized product run on the same CUDA kernels; the (u^T (x) I) Sigma (u (x) I) contraction that assembles the Gram
# entries is small-index glue.  fit() runs Adam on the dense (N n)-dimensional log marginal (mll.dense_log_marginal:
# CUDA Cholesky / inverse / K^-1, the covariance assembled differentiably by HetergeneousCoregionalizationKernel) —
# the reference's speed test fits this comparator for 50 iterations before timing it (pendulum.py:1366).
# =====================================================================================================================
class ControlAffineVectorGP(ControlAffineExactGP):
    def __init__(self, x_dim, u_dim, likelihood, rank=None, gamma_length_scale_prior=None):
        nn.Module.__init__(self)
        self.likelihood = likelihood
        self.matshape = (1 + u_dim, x_dim)
        self.decoder = CatEncoder(1, x_dim, 1 + u_dim)
        self.mean_module = HetergeneousMatrixVariateMean(ConstantMean(), self.decoder, self.matshape)
        num_tasks = int(np.prod(self.matshape))
        self.task_covar = IndexKernel(num_tasks=num_tasks, rank=(num_tasks if rank is None else rank))
        prior = None if gamma_length_scale_prior is None else GammaPrior(*gamma_length_scale_prior)
        self.input_covar = ScaleKernel(RBFKernel(lengthscale_prior=prior) + LinearKernel())
        self.covar_module = HetergeneousCoregionalizationKernel(self.task_covar, self.input_covar, self.decoder)
        self.train_inputs = None
        self.train_targets = None

    def forward(self, mxu):
        return MultivariateNormalResult(self.mean_module(mxu), self.covar_module(mxu, mxu))

    def state_dict(self, *a, **k):
        return dict(matshape=self.matshape, decoder=self.decoder.state_dict(),
                    mean_module=self.mean_module.state_dict(), task_covar=nn.Module.state_dict(self.task_covar),
                    input_covar=nn.Module.state_dict(self.input_covar), train_inputs=self.train_inputs,
                    train_targets=self.train_targets)


class ControlAffineRegressorVector(ControlAffineRegressor):
    def __init__(self, *args, model_class=ControlAffineVectorGP, **kwargs):
        super().__init__(*args, model_class=model_class, **kwargs)

    def _fit_with_warnings(self, Xtrain_in, Utrain_in, XdotTrain_in, training_iter=50, lr=0.1):
        if Xtrain_in.shape[0] == 0:
            return self
        _need_cuda(self.device)
        model = self.model
        Xtrain, Utrain, XdotTrain = [self._ensure_device_dtype(X) for X in (Xtrain_in, Utrain_in, XdotTrain_in)]
        self.clear_cache()
        model.set_train_data(Xtrain, Utrain, XdotTrain)
        model.train()
        optimizer = torch.optim.Adam(model.parameters(), lr=lr)
        scheduler = torch.optim.lr_scheduler.MultiStepLR(
            optimizer, milestones=(torch.tensor([0.3, 0.6, 0.8, 0.90]) * training_iter).tolist())
        MXU = model.train_inputs[0].double()     # the covariance is assembled in float64 whatever the model dtype (the
        # kernels factorise in float64; float32 round-off of a (N n)^2 matrix would need jitter at the 1e-2 level)
        prior = model.input_covar.base_kernel.kernels[0].lengthscale_prior
        self.fit_losses = []
        for i in range(training_iter):          # the loop of the base class (reference :310-334) on the dense density
            optimizer.zero_grad()
            if self._fit_noise_source is not None:
                noise = torch.as_tensor(next(self._fit_noise_source)).reshape(-1).to(device=self.device, dtype=XdotTrain.dtype)
            else:
                noise = torch.rand(XdotTrain.numel(), dtype=XdotTrain.dtype).to(self.device)
            y = (XdotTrain.reshape(-1) * (1 + 1e-6 * noise)).double()
            out = model(MXU)
            logp = dense_log_marginal(out.covariance_matrix, y - out.mean.double())
            if prior is not None:
                logp = logp + prior.log_prob(model.input_covar.base_kernel.kernels[0].lengthscale.double())
            loss = -logp / y.numel()
            assert not torch.isnan(loss).any() and not torch.isinf(loss).any()
            loss.backward()
            self.fit_losses.append(loss.detach())
            optimizer.step()
            scheduler.step()
        return self

    def set_hyperparameters(self, lengthscale=None, outputscale=None, Sigma=None, C=None, linear_variance=None):
        """Constrained hyper-parameters of the CoGP model: scalar RBF lengthscale, outputscale, Sigma (pn,pn), C (p,n),
        LinearKernel variance."""
        from .gp_modules import inv_softplus
        mdl, dev, dt = self.model, self.device, self.dtype
        t = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64) if not isinstance(v, torch.Tensor) else v,
                                      dtype=torch.float64)
        rbf, lin = mdl.input_covar.base_kernel.kernels[0], mdl.input_covar.base_kernel.kernels[1]
        with torch.no_grad():
            if lengthscale is not None:
                rbf.raw_lengthscale.copy_(inv_softplus(t(lengthscale)).reshape(1, 1).to(dev, dt))
            if linear_variance is not None:
                lin.raw_variance.copy_(inv_softplus(t(linear_variance)).reshape(1, 1).to(dev, dt))
            if outputscale is not None:
                mdl.input_covar.raw_outputscale.copy_(inv_softplus(t(outputscale)).reshape(()).to(dev, dt))
            if Sigma is not None:
                M = t(Sigma)
                d = 0.5 * float(torch.linalg.eigvalsh(M).min())
                Fm = torch.linalg.cholesky(M - d * torch.eye(M.shape[0], dtype=torch.float64))
                mdl.task_covar.covar_factor = nn.Parameter(Fm.to(dev, dt))
                mdl.task_covar.raw_var = nn.Parameter(inv_softplus(torch.full((M.shape[0],), d, dtype=torch.float64)).to(dev, dt))
            if C is not None:
                for bm, c in zip(mdl.mean_module.base_means, t(C).reshape(-1)):
                    bm.constant.fill_(float(c))
        self.clear_cache()
        return self

    # ---- pieces -------------------------------------------------------------------------------------------------------
    def _sigma64(self):
        return self.model.task_covar.covar_matrix.evaluate().detach().double().contiguous()

    def _k_data(self, a, c):
        """outputscale * (RBF(a, c) + variance * a c^T) on the GPU (float64)."""
        ic = self.model.input_covar
        rbf, lin = ic.base_kernel.kernels[0], ic.base_kernel.kernels[1]
        s = float(ic.outputscale.detach())
        ls = rbf.lengthscale.detach().reshape(-1).double().expand(a.shape[1]).contiguous()
        K = ops.gram_ca(a.contiguous(), c.contiguous(), ls, s)
        return ops.gemm(a, c, transb=True, alpha=s * float(lin.variance.detach()), beta=1.0, C=K)

    def _u_sigma(self, UH, Sigma):
        """(u_i^T (x) I_n) Sigma  ->  (k, n, p n)."""
        p, n = self.model.matshape
        S4 = Sigma.reshape(p, n, p * n)
        return torch.einsum('iq,qrc->irc', UH, S4)

    def _perturbed_cholesky_compute(self, k_xx, Sigma, Xtrain, UHtrain, cholesky_tries=10, cholesky_perturb_init=1e-5,
                                    cholesky_perturb_scale=10):
        _need_cuda(Xtrain.device)
        p, n = self.model.matshape
        X64, UH64 = Xtrain.double().contiguous(), UHtrain.double().contiguous()
        k = X64.shape[0]
        KXX = self._k_data(X64, X64)
        US = self._u_sigma(UH64, Sigma)                                            # (k, n, p n)
        uSu = torch.einsum('irqs,jq->irjs', US.reshape(k, n, p, n), UH64)            # (k, n, k, n)
        Kb = (KXX.reshape(k, 1, k, 1) * uSu).reshape(k * n, k * n)
        Kbp, L = self._make_psd_like_reference(Kb, cholesky_tries, cholesky_perturb_init, cholesky_perturb_scale)
        return L

    def _make_psd_like_reference(self, M, tries=10, init=1e-5, scale=10, want_inverse=True):
        nrows = M.shape[0]
        factor = init
        for ntry in range(tries):
            eps = next(self._jitter_source) if self._jitter_source is not None else _draw_jitter(nrows, self.dtype)
            eps = eps.to(device=M.device, dtype=torch.float64).contiguous()
            buf = torch.eye(ops.padded(nrows), dtype=torch.float64, device=M.device)
            buf[:nrows, :nrows] = M
            try:
                L, dinv = ops.potrf_(buf, nrows, eps, factor)
                break
            except RuntimeError as e:
                if ntry == tries - 1:
                    raise
                LOG.warning("Cholesky failed with perturb={} on error {}".format(factor, str(e)))
                factor *= scale
        if want_inverse:
            self._cache['_Lpad'] = L
            self._cache['_Linv'] = ops.trtri(L, dinv)
        return M + factor * torch.diag(eps), L[:nrows, :nrows]

    def _custom_predict_matrix(self, Xtest_in, Xtestp_in=None, compute_cov=True, _out_jitter=True):
        """M_k (b,n,p) and Sigma_k (b,b',pn,pn)  (reference :1219-1331)."""
        _need_cuda(self.device)
        Xtest = self._ensure_device_dtype(Xtest_in)
        Xtestp = self._ensure_device_dtype(Xtestp_in) if Xtestp_in is not None else Xtest
        out_dt = self.dtype
        p, n = self.model.matshape
        Sigma = self._sigma64()
        C = self.model.mean_module.constants().detach().double()
        Xq, Xp = Xtest.double(), Xtestp.double()
        b, bp_ = Xq.shape[0], Xp.shape[0]
        M0 = C.t().unsqueeze(0).expand(b, n, p)
        if self.model.train_inputs is None:
            return M0.to(out_dt), (Sigma * self._k_data(Xq, Xp).unsqueeze(-1).unsqueeze(-1)).to(out_dt)
        Xtrain, UHtrain, targets = self._train_data()
        X64, UH64 = Xtrain.double().contiguous(), UHtrain.double().contiguous()
        k = X64.shape[0]
        self._perturbed_cholesky(None, Sigma, Xtrain, UHtrain)
        Linv = self._cache['_Linv']
        Npad = Linv.shape[0]
        Y = (targets.double() - UH64 @ C).reshape(-1, 1)                           # vec, (k n, 1), r fastest
        if '_alpha_vec' not in self._cache:
            Ypad = torch.zeros(Npad, 2, dtype=torch.float64, device=Y.device)
            Ypad[:k * n, :1] = Y
            z = ops.trmm_lower(Linv, Ypad)
            self._cache['_alpha_vec'] = ops.trmm_lower(Linv, z.contiguous(), trans=True)[:, :1].contiguous()
        alpha = self._cache['_alpha_vec']
        US = self._u_sigma(UH64, Sigma)                                            # (k, n, pn)
        Kst = self._k_data(Xq, X64)                                                # (b, k)
        # kb_star as an (Npad, b * pn) matrix: rows (i, r), columns (t, c)
        kbs = (Kst.t().reshape(k, 1, b, 1) * US.reshape(k, n, 1, p * n)).reshape(k * n, b * p * n)
        kbs_pad = torch.zeros(Npad, b * p * n, dtype=torch.float64, device=kbs.device)
        kbs_pad[:k * n] = kbs
        mean_k = M0 + ops.gemm(kbs_pad, alpha, transa=True).reshape(b, p, n).transpose(-2, -1)
        if not compute_cov:
            return mean_k.to(out_dt), Xtest.new_zeros(b, bp_, p * n, p * n)
        V = ops.trmm_lower(Linv, kbs_pad)
        KS = torch_kron(self._k_data(Xq, Xp), Sigma, batch_dims=0)
        KkXX = ops.gemm(V, V, transa=True, alpha=-1.0, beta=1.0, C=KS)
        if _out_jitter:
            KkXX, _ = self._make_psd_like_reference(KkXX, want_inverse=False)
        KkXX = KkXX.reshape(b, p * n, bp_, p * n).transpose(2, 1)
        return mean_k.to(out_dt), KkXX.to(out_dt)

    def custom_predict(self, Xtest_in, Utest_in=None, UHfill=1, Xtestp_in=None, Utestp_in=None, UHfillp=1,
                       compute_cov=True):
        Xtest = self._ensure_device_dtype(Xtest_in)
        Xtestp = self._ensure_device_dtype(Xtestp_in) if Xtestp_in is not None else Xtest
        meanFX, KkXX = self._custom_predict_matrix(Xtest_in, Xtestp_in, compute_cov=compute_cov)
        UHtest = self._uh(Xtest, Utest_in, UHfill)
        meanFXU = meanFX.bmm(UHtest.unsqueeze(-1)).squeeze(-1)
        b, n = Xtest.shape
        if compute_cov:
            # the reference contracts BOTH sides with UHtest (:1160-1168), also when Utestp is given
            In = torch.eye(n, dtype=Xtest.dtype, device=Xtest.device)
            blk = torch_kron(UHtest, In, batch_dims=0).reshape(b, 1, n, -1)
            blk_T = blk.reshape(1, b, n, -1).transpose(-2, -1)
            varFXU = torch.matmul(torch.matmul(blk, KkXX), blk_T)
        else:
            varFXU = Xtest.new_zeros(b, Xtestp.shape[0], n, n)
        return meanFXU, varFXU

    def custom_predict_fullmat(self, Xtest_in, Xtestp_in=None):
        Xtest = self._ensure_device_dtype(Xtest_in)
        meanFX, varFX = self._custom_predict_matrix(Xtest_in, Xtestp_in, compute_cov=True)
        b, p, n = Xtest.shape[0], 1 + self.u_dim, self.x_dim
        return meanFX.transpose(-2, -1).reshape(-1), varFX.transpose(2, 1).reshape(b * p * n, b * p * n)

    def custom_predict_blocks(self, *a, **k):
        raise NotImplementedError("per-query block fast path exists for the matrix-variate model only")

    def predict(self, Xtest_in, return_cov=True):
        Xtest = self._ensure_device_dtype(Xtest_in)
        if isinstance(Xtest_in, np.ndarray):
            Xtest_in = torch.from_numpy(Xtest_in)
        meanFX, varFX = self._custom_predict_matrix(Xtest, None, compute_cov=return_cov, _out_jitter=False)
        b, p, n = Xtest.shape[0], 1 + self.u_dim, self.x_dim
        mean = meanFX.transpose(-2, -1).to(device=Xtest_in.device, dtype=Xtest_in.dtype)
        if not return_cov:
            return mean
        cov = varFX.transpose(2, 1).reshape(b * p * n, b * p * n)
        return mean, cov.to(device=Xtest_in.device, dtype=Xtest_in.dtype)


ControlAffineRegVectorDiag = partial(ControlAffineRegressorVector, model_class=partial(ControlAffineVectorGP, rank=0))
