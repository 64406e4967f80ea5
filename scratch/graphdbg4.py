import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch as th
from oracle import cpu_oracle as O
from test_gpu_runner import make_args, build
rng = np.random.default_rng(37)
B, n, m, T = 6, 10, 12, 6
S = O.gen_dense(rng, B, n, m, T)
env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=3, M=4, L=3, lambda_=0.5, sat_prox_mat=S, graphs=1)
def roll(use_graph, noise, eps=8):
    args = make_args("real_constellation_env", env_args, B, reuse_episode_batch=True, use_cuda_graph=use_graph, epsilon_start=0.5, epsilon_finish=0.5)
    runner, mac, buffer, logger = build(args)
    res=[]
    for ep in range(eps):
        if noise: junk = th.randn(1<<22, device='cuda').sort()[0]
        batch = runner.run()
        res.append((batch["actions"].clone(), batch["obs"].clone()))
    return res
def cmp(a,b): return [bool(th.equal(x[0],y[0]) and th.equal(x[1],y[1])) for x,y in zip(a,b)]
e1=roll(False,False); e2=roll(False,True); g1=roll(True,False); g2=roll(True,True)
print("eager vs eager+noise", cmp(e1,e2))
print("graph vs graph+noise", cmp(g1,g2))
print("eager vs graph      ", cmp(e1,g1))
