import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch as th
from oracle import cpu_oracle as O
from test_gpu_runner import make_args, build
rng = np.random.default_rng(37)
B, n, m, T = 6, 10, 12, 6
S = O.gen_dense(rng, B, n, m, T)
env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=3, M=4, L=3, lambda_=0.5, sat_prox_mat=S, graphs=1)
for name, kw, test_mode in [("const eps", dict(epsilon_start=0.5, epsilon_finish=0.5), False), ("greedy", {}, True), ("anneal", dict(epsilon_anneal_time=200), False)]:
    outs=[]
    for use_graph in (False, True):
        args = make_args("real_constellation_env", env_args, B, reuse_episode_batch=True, use_cuda_graph=use_graph, **kw)
        runner, mac, buffer, logger = build(args)
        res=[]
        for ep in range(4):
            batch = runner.run(test_mode=test_mode)
            res.append(batch["actions"].clone())
            if use_graph and ep>=1: print(name, ep, "eps_dev", mac.action_selector._eps_dev.item(), "eps", mac.action_selector.epsilon)
        outs.append(res)
    print(name, [bool(th.equal(a,b)) for a,b in zip(*outs)], [int((a!=b).sum()) for a,b in zip(*outs)])
