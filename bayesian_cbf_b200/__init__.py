"""bayesian_cbf_b200 — B200-native matrix-variate GP hot path of Bayesian_CBF.

Host side: the reference's Python API (ControlAffineRegressor & co).  Device side: libbcbf.so
(hand-written sm_100a CUDA behind the C ABI of include/bcbf.h).  No CPU fallback."""
__version__ = '0.1.0'
