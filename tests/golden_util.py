"""Helpers for reading the reference-generated golden fixtures (tests/golden/*.npz)."""
import os

import numpy as np
import torch

from oracle.mvgp_oracle import Hyper

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

PREDICT_CASES = ['ref_predict_unicycle_f64', 'ref_predict_pendulum_f64', 'ref_predict_pendulum_f32',
                 'ref_predict_unicycle_rank1_fit_f64']


def load(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def T(a):
    return torch.from_numpy(np.asarray(a, dtype=np.float64))


def hyper(d, prefix):
    return Hyper(lengthscale=T(d[prefix + 'lengthscale']), outputscale=T(d[prefix + 'outputscale']),
                 A=T(d[prefix + 'A']), B=T(d[prefix + 'B']), C=T(d[prefix + 'C']))


def rands(d, tag):
    return [T(d['%s_rand%d' % (tag, i)]) for i in range(int(d[tag + '_nrand']))]


def tol(d):
    """Goldens were produced in the case's dtype; float32 cases are compared at float32 resolution."""
    return dict(rtol=1e-9, atol=1e-11) if str(d['dtype']) == 'float64' else dict(rtol=2e-3, atol=2e-4)
