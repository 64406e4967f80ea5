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
def addmm_act():
    h=th._addmm_activation(b1,x,W1t); h2=th._addmm_activation(b2,h,W2t); return th.addmm(b3,h2,W3t)
t(addmm_act,"_addmm_activation (fused relu)")
print("max diff fused", (addmm()-addmm_act()).abs().max().item())
W3p=th.zeros(64,128,device='cuda'); W3p[:,:100]=W3t; b3p=th.zeros(128,device='cuda'); b3p[:100]=b3
def addmm_pad():
    h=th._addmm_activation(b1,x,W1t); h2=th._addmm_activation(b2,h,W2t); return th.addmm(b3p,h2,W3p)
t(addmm_pad,"fused relu + out padded to 128 cols")
t(lambda: th.addmm(b3,x[:, :64].contiguous() if False else th.relu(th.addmm(b2, th.relu(th.addmm(b1,x,W1t)), W2t)), W3t),"addmm + th.relu")
h2_=th.relu(th.addmm(b2, th.relu(th.addmm(b1,x,W1t)), W2t))
t(lambda: th.addmm(b3,h2_,W3t),"out layer only (64->100)")
t(lambda: th.addmm(b3p,h2_,W3p),"out layer only padded (64->128)")
h1_=th.relu(th.addmm(b1,x,W1t))
t(lambda: th.addmm(b2,h1_,W2t),"fc2 only (64->64)")
t(lambda: th.addmm(b1,x,W1t),"fc1 only (490->64)")
print("---- per layer")
t(lambda: th._addmm_activation(b1,x,W1t),"fc1 _addmm_activation")
def fc1_mm_inplace():
    h=th.mm(x,W1t); h.add_(b1); h.relu_(); return h
t(fc1_mm_inplace,"fc1 mm + add_ + relu_")
hbuf=th.empty(R,64,device='cuda')
def fc1_mm_out():
    th.mm(x,W1t,out=hbuf); hbuf.add_(b1).relu_(); return hbuf
t(fc1_mm_out,"fc1 mm(out=) + add_ + relu_")
t(lambda: th._addmm_activation(b2,h1_,W2t),"fc2 _addmm_activation")
t(lambda: hbuf.add_(b1),"add_ bias only [R,64]")
t(lambda: hbuf.relu_(),"relu_ only [R,64]")
print("---- mm only, narrow layers")
t(lambda: th.mm(h1_,W2t),"fc2 mm only (64->64)")
t(lambda: th.mm(h2_,W3t),"out mm only (64->100)")
W3t128=th.zeros(64,128,device='cuda'); W3t128[:,:100]=W3t
t(lambda: th.mm(h2_,W3t128),"out mm only padded (64->128)")
t(lambda: F.linear(h2_,W3,b3),"out F.linear (64->100)")
t(lambda: th.matmul(h2_,W3t),"out matmul (64->100)")
h2h=h2_.half(); W3h=W3t.half()
t(lambda: th.mm(h2h,W3h),"out mm fp16 (reference only)")
