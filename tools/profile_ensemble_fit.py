import sys, time, torch
sys.path.insert(0, '/root/repo')
from bayesian_cbf_b200.ensemble import EnsembleHyperParameters, fit_ensemble_hyperparameters
R, N = 512, 200
g = torch.Generator().manual_seed(0)
X = torch.zeros(R, N, 3, dtype=torch.float64); X[:, :, 2] = 6.28 * torch.rand(R, N, generator=g, dtype=torch.float64)
U = 2 * torch.rand(R, N, 2, generator=g, dtype=torch.float64) - 1
Xdot = torch.sin(X) * (1 + U[:, :, :1]) + 0.01 * torch.randn(R, N, 3, generator=g, dtype=torch.float64)
X, U, Xdot = X.cuda(), U.cuda(), Xdot.cuda()
hp = EnsembleHyperParameters(R, 3, 3, rank=1, device='cuda')
fit_ensemble_hyperparameters(hp, X, U, Xdot, training_iter=5, lr=0.1, generator=g)
torch.cuda.synchronize()
t0 = time.perf_counter()
fit_ensemble_hyperparameters(hp, X, U, Xdot, training_iter=20, lr=0.1, generator=g)
torch.cuda.synchronize()
print('ms per iteration', 1e3 * (time.perf_counter() - t0) / 20)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    fit_ensemble_hyperparameters(hp, X, U, Xdot, training_iter=10, lr=0.1, generator=g)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
