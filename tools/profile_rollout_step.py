import math, sys, time, torch
sys.path.insert(0,'/root/repo')
from bayesian_cbf_b200 import unicycle as U, ops
R, dt = 512, 0.001
x0=[-3.0,-1.0,-math.pi/4]; xg=[0.0,0.0,math.pi/4]
planner=U.PiecewiseLinearPlanner(x0,xg,2000,dt,frac_time_to_reach_goal=0.95)
cbfs=U.obstacles_at_mid_from_start_and_goal(x0,xg,term_weights=(0.7,0.3))
ctrl=U.BayesCBFController(planner,U.CLFCartesian(Kp=(0.9,1.5,0.0)),cbfs,[5.0,5.0],model_L=12.0,clf_gamma=10.0,max_risk=0.01)
g=torch.Generator().manual_seed(0)
X=(torch.tensor(x0,dtype=torch.float64).repeat(R,1)+0.05*(torch.rand(R,3,generator=g,dtype=torch.float64)-0.5)).cuda()
def tm(f,n=20):
    f(); torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(n): r=f()
    torch.cuda.synchronize(); return (time.perf_counter()-t)/n*1e3, r
ms,(c,d,A,b)=tm(lambda: ctrl.constraint_terms(X,3)); print('constraint_terms ms',ms)
w=torch.tensor([0.33,0.33,0.33],dtype=torch.float64,device='cuda')
ms,(y,st,it)=tm(lambda: ops.socp_solve(w,c.contiguous(),d.contiguous(),A.contiguous(),b.contiguous(),ctrl.rho)); print('socp_solve ms',ms,'iters mean',it.float().mean().item(),'max',it.max().item(),'infeasible',int(st.sum()))
ms,_=tm(lambda: ctrl.control(X,3)); print('control ms',ms)
