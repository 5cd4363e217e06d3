import torch, sys
sys.path.insert(0, '/root/repo')
from bayesian_cbf_b200 import ops
g = torch.Generator().manual_seed(0)
for n in (128, 256, 1024):
    R = torch.randn(n, n, generator=g, dtype=torch.float64)
    A0 = (R @ R.T + n * torch.eye(n, dtype=torch.float64)).cuda()
    ts = []
    for rep in range(10):
        A = A0.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.potrf_(A, n, None, 0.0, check_pd=False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print('potrf n=%d: min %.1f us' % (n, 1e3 * min(ts)))
