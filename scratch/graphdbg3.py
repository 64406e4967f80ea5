import torch as th
th.manual_seed(0)
x=th.randn(600,70,device='cuda'); W=th.randn(70,64,device='cuda'); b=th.randn(64,device='cuda')
W2=th.randn(64,12,device='cuda'); b2=th.randn(12,device='cuda')
def f():
    h=th.relu(th.addmm(b,x,W)); return th.addmm(b2,h,W2)
e=f().clone()
g=th.cuda.CUDAGraph()
th.cuda.synchronize()
with th.cuda.graph(g):
    out=f()
g.replay(); th.cuda.synchronize()
print("graph vs eager equal:", th.equal(out,e), (out-e).abs().max().item())
g.replay(); th.cuda.synchronize(); o2=out.clone(); g.replay(); th.cuda.synchronize()
print("graph self-consistent:", th.equal(o2,out))
