"""Mean module of the MVGP (reference bayes_cbf/matrix_variate_multitask_model.py:9-76): p*n independent
ConstantMeans arranged as C (p, n).  Train rows (mask 1) give uh_i^T C (n values), test rows (mask 0) give vec(C)."""
import torch

from .gp_modules import MultitaskMean
from .matrix_variate_multitask_kernel import prod


class HetergeneousMatrixVariateMean(MultitaskMean):
    def __init__(self, mean_module, decoder, matshape, **kwargs):
        super().__init__(mean_module, prod(matshape), **kwargs)
        self.decoder = decoder
        self.matshape = matshape

    def constants(self):
        """C (p, n): C[q, r] = base_means[q*n + r].constant."""
        return torch.stack([bm.constant.reshape(()) for bm in self.base_means]).reshape(*self.matshape)

    def mean1(self, UH, mu):
        return (UH @ mu).reshape(-1) if mu.ndim == 2 else (UH.unsqueeze(-2) @ mu).reshape(-1)

    def mean2(self, mu):
        return mu.reshape(-1)

    def forward(self, MXU):
        """Flat mean of a train-first sorted MXU: uh_i^T C (n values) per train row, then vec(C) (p*n) per test row."""
        assert not torch.isnan(MXU).any()
        Ms, _, UH = self.decoder.decode(MXU)
        assert Ms.size(-1) == 1
        from .matrix_variate_multitask_kernel import _train_end
        mask = Ms[..., 0]
        ntrain, ntotal = _train_end(mask), mask.size(-1)
        assert (mask[ntrain:] == 0).all(), "rows must be sorted train-first"
        C = self.constants().to(MXU.dtype)
        pieces = []
        if ntrain:
            pieces.append(self.mean1(UH[:ntrain], C))
        if ntotal > ntrain:
            pieces.append(self.mean2(C.unsqueeze(0).expand(ntotal - ntrain, *self.matshape)))
        return torch.cat(pieces)

    def state_dict(self, *a, **k):
        return dict(matshape=self.matshape, decoder=self.decoder.state_dict(),
                    constants=self.constants().detach().clone())

    def load_state_dict(self, state_dict, *a, **k):
        self.matshape = state_dict['matshape']
        self.decoder.load_state_dict(state_dict['decoder'])
        if 'constants' in state_dict:
            with torch.no_grad():
                for bm, c in zip(self.base_means, state_dict['constants'].reshape(-1)):
                    bm.constant.fill_(float(c))
