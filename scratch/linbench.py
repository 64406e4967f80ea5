import os, sys, time, torch as th
import torch.nn.functional as F
th.manual_seed(0)
R=409600
x=th.randn(R,490,device='cuda'); 
W1=th.randn(64,490,device='cuda')*0.05; b1=th.randn(64,device='cuda')
W2=th.randn(64,64,device='cuda')*0.1; b2=th.randn(64,device='cuda')
W3=th.randn(100,64,device='cuda')*0.1; b3=th.randn(100,device='cuda')
def t(fn,name,n=20):
    for _ in range(3): fn()
    th.cuda.synchronize(); e0=th.cuda.Event(enable_timing=True); e1=th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); th.cuda.synchronize()
    print(f"{name:40s} {e0.elapsed_time(e1)/n*1000:9.1f} us")
def lin():
    h=F.relu(F.linear(x,W1,b1)); h=F.relu(F.linear(h,W2,b2)); return F.linear(h,W3,b3)
def mm_addrelu():
    h=th.mm(x,W1.t()); h.add_(b1).relu_(); h2=th.mm(h,W2.t()); h2.add_(b2).relu_(); q=th.mm(h2,W3.t()); q.add_(b3); return q
W1t=W1.t().contiguous(); W2t=W2.t().contiguous(); W3t=W3.t().contiguous()
def mm_pre():
    h=th.mm(x,W1t); h.add_(b1).relu_(); h2=th.mm(h,W2t); h2.add_(b2).relu_(); q=th.mm(h2,W3t); q.add_(b3); return q
def addmm():
    h=th.addmm(b1,x,W1t).relu_(); h2=th.addmm(b2,h,W2t).relu_(); return th.addmm(b3,h2,W3t)
t(lin,"F.linear x3 (+relu)")
t(mm_addrelu,"mm + add_ + relu_")
t(mm_pre,"mm (pre-transposed W) + add_ + relu_")
t(addmm,"addmm")
t(lambda: F.linear(x,W1,b1),"fc1 F.linear only")
t(lambda: th.mm(x,W1t),"fc1 mm only")
t(lambda: F.linear(x,W1),"fc1 F.linear nobias")
print("max diff", (lin()-mm_pre()).abs().max().item())
