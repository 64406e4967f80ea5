import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch as th
from oracle import cpu_oracle as O
from test_gpu_runner import make_args, build
rng = np.random.default_rng(37)
B, n, m, T = 6, 10, 12, 6
S = O.gen_dense(rng, B, n, m, T)
env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=3, M=4, L=3, lambda_=0.5, sat_prox_mat=S, graphs=1)
outs=[]
for use_graph in (False, True):
    args = make_args("real_constellation_env", env_args, B, epsilon_anneal_time=200, reuse_episode_batch=True, use_cuda_graph=use_graph)
    runner, mac, buffer, logger = build(args)
    res=[]
    for ep in range(4):
        batch = runner.run()
        res.append({k: v.clone() for k,v in batch.data.transition_data.items()})
        print(use_graph, ep, "eps", mac.action_selector.epsilon, "ctr", runner.episode_ctr.item(), "k", runner.env.k[:3].tolist(), "ret", runner.last_episode_returns[:2].tolist())
    outs.append(res)
for ep in range(4):
    for k in outs[0][ep]:
        a,b=outs[0][ep][k],outs[1][ep][k]
        if not th.equal(a,b):
            d=(a!=b)
            idx=d.nonzero()[0].tolist()
            print("ep",ep,"field",k,"first diff idx",idx, "count", int(d.sum()), a[tuple(idx)].item(), b[tuple(idx)].item())
