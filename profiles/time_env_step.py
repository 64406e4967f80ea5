"""Time the real-env step (reset + T steps, CUDA events per step) at any shape on a chosen kernel path.
    python profiles/time_env_step.py B n m [path] [reps]      (path: sap_real_select_kernel code, 0 = automatic)"""
import sys

import torch as th

from marl_sap_b200 import _lib
from marl_sap_b200.components.episode_buffer import EpisodeBatch
from marl_sap_b200.envs.batched import BatchedRealConstellationEnv, real_obs_size, real_scheme

B, n, m = (int(x) for x in sys.argv[1:4])
path = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
T = 8
g = th.Generator().manual_seed(0)
S = th.rand(B, n, m, T, generator=g).cuda()
assert _lib.load().sap_real_select_kernel(path) >= 0
env = BatchedRealConstellationEnv(B, n, m, T, 3, 10, 10, 0.5, sat_prox_mat=S)
scheme, pre = real_scheme(n, m, 3, real_obs_size(10, 10, 3))
batch = EpisodeBatch(scheme, {"agents": n}, B, T + 1, preprocess=pre, device="cuda", lazy=("beta", "avail_actions", "actions_onehot"))
batch.agent_in = th.zeros(B, n, env.obs_size, device="cuda")
acts = [th.randint(0, m, (B, n), generator=g).cuda() for _ in range(T)]
best = []
for r in range(reps):
    env.reset(batch)
    evs = []
    for t in range(T - 3):
        a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        a.record()
        env.step(acts[t], batch)
        b.record()
        evs.append((a, b))
    th.cuda.synchronize()
    best.append(sum(a.elapsed_time(b) for a, b in evs[1:]) / (len(evs) - 1))
print(f"B={B} n={n} m={m} path={path}: env step {min(best):.4f} ms (min over {reps} episodes of the mean over {T - 4} steps)")
