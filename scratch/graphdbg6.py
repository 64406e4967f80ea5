import sys, os; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch as th
from oracle import cpu_oracle as O
from test_gpu_runner import make_args, build
rng = np.random.default_rng(37)
B, n, m, T = 6, 10, 12, 6
S = O.gen_dense(rng, B, n, m, T)
def roll(use_graph, eps=10, **kw):
    env_args = dict(n=n, m=m, T=T, L=3, lambda_=0.5, sat_prox_mat=S); rk={"prev0": np.tile(np.arange(n),(B,1))}
    args = make_args("mock_constellation_env", env_args, B, reuse_episode_batch=True, use_cuda_graph=use_graph, epsilon_start=0.5, epsilon_finish=0.5, **kw)
    runner, mac, buffer, logger = build(args)
    res=[]
    for ep in range(eps):
        batch = runner.run(**rk)
        th.cuda.synchronize()
        res.append({k: v.clone() for k,v in batch.data.transition_data.items()})
        res[-1]["ctr"]=runner.episode_ctr.clone(); res[-1]["k"]=runner.env.k.clone()
    return res
a=roll(False); b=roll(True)
for ep in range(10):
    bad=[k for k in a[ep] if not th.equal(a[ep][k], b[ep][k])]
    msg=""
    if "actions" in bad:
        d=(a[ep]["actions"]!=b[ep]["actions"]).nonzero()
        msg=f" first action diff (b,t,i)={d[0].tolist()[:3]} n={len(d)} ts={sorted(set(d[:,1].tolist()))}"
    print(ep, bad, msg, "ctr", a[ep]["ctr"].item(), b[ep]["ctr"].item())
