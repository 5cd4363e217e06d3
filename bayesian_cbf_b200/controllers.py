"""The consumer side of the hot path: CBC terms -> second-order-cone terms (reference
bayes_cbf/controllers.py:423-482 `SOCPController.convert_cbc_terms_to_socp_terms`, stricter twin at
bayes_cbf/unicycle_move_to_pose.py:837-878).  The solver orchestration of the reference's controllers (cvxpy / GUROBI)
is out of scope."""
import torch

from . import ops


def _compute_device(dev):
    if dev.type == 'cuda':
        return dev
    if not torch.cuda.is_available():
        raise RuntimeError("convert_cbc_terms_to_socp_terms factorises on a CUDA device (no CPU fallback)")
    return torch.device('cuda', torch.cuda.current_device())


def convert_cbc_terms_to_socp_terms(bfe, e, V, bfv, v, extravars, testing=False, singular_fallback=True):
    """mean(u) = bfe^T u + e,  var(u) = u^T V u + bfv^T u + v   ->   (A, bfb, bfc, d) with
        || A y + bfb ||_2 <= bfc^T y + d,   y = [extra vars (the last one is the relaxation delta); u],
    where Asq = [[v, bfv^T/2],[bfv/2, V]] = L L^T, A = [0 | L^T[:,1:]], bfb = L^T[:,0].
    The factorisation runs on the GPU (bcbf_socp_factor); `singular_fallback` retries with Asq + 1e-3 I like
    controllers.py:447-449 and otherwise a non-PD Asq raises RuntimeError like torch.cholesky does there."""
    assert extravars >= 1, "I assumed atleast δ "
    m = bfe.shape[-1]
    dt, dev = bfe.dtype, bfe.device
    with torch.no_grad():
        v_ = torch.as_tensor(v, dtype=dt, device=dev).reshape(1, 1)
        Asq = torch.cat((torch.cat((v_, (bfv / 2).reshape(1, -1)), dim=-1),
                         torch.cat(((bfv / 2).reshape(-1, 1), V), dim=-1)), dim=-2)
        cdev = _compute_device(dev)
        A_s, b_s, status = ops.socp_factor(Asq.to(device=cdev, dtype=torch.float64).reshape(1, m + 1, m + 1).contiguous(),
                                           reg=1e-3 if singular_fallback else 0.0)
        if int(status.item()) != 0:
            raise RuntimeError("cholesky: Asq is not positive-definite (pivot %d)" % int(status.item()))
        A = torch.zeros((m + 1, m + extravars), dtype=dt, device=dev)
        A[:, extravars:] = A_s[0].to(device=dev, dtype=dt)
        bfb = b_s[0].to(device=dev, dtype=dt)
    bfc = bfe.new_zeros((m + extravars))
    bfc[extravars - 1] = 1      # the relaxation variable
    bfc[extravars:] = bfe
    return A, bfb, bfc, e
