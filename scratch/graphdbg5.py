import sys, os; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch as th
from oracle import cpu_oracle as O
from test_gpu_runner import make_args, build
rng = np.random.default_rng(37)
B, n, m, T = 6, 10, 12, 6
S = O.gen_dense(rng, B, n, m, T)
def roll(env_name, use_graph, eps=12, **kw):
    if env_name=="real_constellation_env":
        env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=3, M=4, L=3, lambda_=0.5, sat_prox_mat=S, graphs=1); rk={}
    else:
        env_args = dict(n=n, m=m, T=T, L=3, lambda_=0.5, sat_prox_mat=S); rk={"prev0": np.tile(np.arange(n),(B,1))}
    args = make_args(env_name, env_args, B, reuse_episode_batch=True, use_cuda_graph=use_graph, epsilon_start=0.5, epsilon_finish=0.5, **kw)
    runner, mac, buffer, logger = build(args)
    res=[]
    for ep in range(eps):
        batch = runner.run(**rk)
        res.append((batch["actions"].clone(), batch["obs"].clone()))
    return res
def cmp(a,b): return "".join("1" if (th.equal(x[0],y[0]) and th.equal(x[1],y[1])) else "0" for x,y in zip(a,b))
print("mock               ", cmp(roll("mock_constellation_env",False), roll("mock_constellation_env",True)))
print("real fast          ", cmp(roll("real_constellation_env",False), roll("real_constellation_env",True)))
print("real fast no-stage ", cmp(roll("real_constellation_env",False,stage_agent_inputs=False), roll("real_constellation_env",True,stage_agent_inputs=False)))
os.environ["SAP_REAL_FORCE_GENERIC"]="1"
print("real generic       ", cmp(roll("real_constellation_env",False), roll("real_constellation_env",True)))
