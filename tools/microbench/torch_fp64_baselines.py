"""Library FP64 baselines on the box (cuBLAS DGEMM / cuSOLVER potrf / trsm) — roofline denominators
and 'what stock torch does' reference points. Not part of the product path."""
import torch, time, json
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best=1e9
    for _ in range(reps):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best=min(best,e0.elapsed_time(e1))
    return best
out={}
d='cuda'
for n in (4096,8192):
    a=torch.randn(n,n,device=d,dtype=torch.float64); b=torch.randn(n,n,device=d,dtype=torch.float64)
    ms=t(lambda: a@b); out[f'dgemm_{n}_tflops']=2*n**3/ms*1e-9
n=16384
a=torch.randn(n,n,device=d,dtype=torch.float64)
k=a@a.T/n+torch.eye(n,device=d,dtype=torch.float64)
ms=t(lambda: torch.linalg.cholesky(k),2); out['potrf_16384_ms']=ms; out['potrf_16384_tflops']=n**3/3/ms*1e-9
L=torch.linalg.cholesky(k)
r=torch.randn(n,4096,device=d,dtype=torch.float64)
ms=t(lambda: torch.linalg.solve_triangular(L,r,upper=False),2); out['trsm_16384x4096_ms']=ms; out['trsm_tflops']=n*n*4096/ms*1e-9
ms=t(lambda: torch.linalg.solve_triangular(L,torch.eye(n,device=d,dtype=torch.float64),upper=False),1); out['trtri_via_trsm_16384_ms']=ms
Lt=torch.tril(L)
ms=t(lambda: Lt@r,2); out['gemm_16384x16384x4096_tflops']=2*n*n*4096/ms*1e-9
print(json.dumps(out,indent=1))
