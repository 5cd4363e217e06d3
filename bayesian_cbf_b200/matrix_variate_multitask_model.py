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
        return (UH.unsqueeze(-2) @ mu).reshape(-1)

    def mean2(self, mu):
        return mu.reshape(-1)

    def forward(self, MXU):
        assert not torch.isnan(MXU).any()
        Ms, _, UH = self.decoder.decode(MXU)
        assert Ms.size(-1) == 1
        Ms = Ms[..., 0]
        idxs = torch.nonzero(Ms - Ms.new_ones(Ms.size()))
        idxend = int(torch.min(idxs)) if idxs.numel() else Ms.size(-1)
        mu = self.constants().to(MXU.dtype).unsqueeze(0).expand(Ms.size(-1), *self.matshape)
        output = None
        if idxend != 0:
            assert (Ms[..., idxend:] == 0).all()
            output = self.mean1(UH[..., :idxend, :], mu[:idxend, ...])
        if Ms.size(-1) != idxend:
            Fmean = self.mean2(mu[idxend:, ...])
            output = torch.cat([output, Fmean]) if output is not None else Fmean
        return output

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
