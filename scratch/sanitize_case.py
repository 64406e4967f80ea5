"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as th
from types import SimpleNamespace
from marl_sap_b200.envs.batched import BatchedRealConstellationEnv, BatchedMockConstellationEnv
from marl_sap_b200.components.episode_buffer import EpisodeBatch, ReplayBuffer
from marl_sap_b200.action_selectors import REGISTRY as SEL
rng = np.random.default_rng(0)
def batch_for(env, B, lazy=()):
    b = EpisodeBatch(copy.deepcopy(env.scheme), {"agents": env.n}, B, env.T + 1, preprocess=env.preprocess, device="cuda", lazy=lazy)
    b.agent_in = th.zeros(B, env.n, env.obs_size, device="cuda")
    return b
args = SimpleNamespace(epsilon_start=0.5, epsilon_finish=0.5, epsilon_anneal_time=1, evaluation_epsilon=0.0, seed=1, env_args={"M": 10, "m": 52})
for (B, n, m, T, L, M, N, prios, ties) in [(3, 50, 52, 3, 3, 10, 10, False, False), (2, 20, 33, 3, 2, 6, 5, True, True), (2, 100, 100, 2, 3, 10, 10, False, True)]:
    S = rng.random((B, n, m, T), dtype=np.float32)
    if ties:
        S = (np.round(S * 4) / 4).astype(np.float32)
    pr = (rng.integers(1, 4, size=m) * 0.5).astype(np.float32) if prios else None
    env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5, sat_prox_mat=S, task_prios=pr)
    batch = batch_for(env, B)
    sel = SEL["epsilon_greedy"](args)
    env.reset(batch)
    for t in range(T):
        q = th.randn(B, n, m, device="cuda")
        a = sel.select_action(q, batch["avail_actions"][:, t], 0)
        env.step(a, batch)
    rb = ReplayBuffer(copy.deepcopy(env.scheme), {"agents": n}, B + 1, T + 1, preprocess=env.preprocess, device="cuda")
    rb.insert_episode_batch(batch); rb.insert_episode_batch(batch)
    rb.gather(np.array([0, 1]))
    fs = SEL["filtered_const_epsilon_greedy"](SimpleNamespace(**{**args.__dict__, "env_args": {"M": M, "m": m}}))
    fs.select_action(th.randn(B, n, M + 1, device="cuda"), batch["avail_actions"][:, 0], 0, top=env.top)
    fs.select_action(th.randn(B, n, M + 1, device="cuda"), batch["avail_actions"][:, 0], 0, beta=batch["beta"][:, 0])
env = BatchedMockConstellationEnv(3, 10, 12, 4, 3, 0.5, sat_prox_mat=rng.random((3, 10, 12, 4), dtype=np.float32))
batch = batch_for(env, 3)
env.reset(batch, prev0=np.tile(np.arange(10), (3, 1)))
for t in range(4):
    env.step(th.randint(0, 12, (3, 10), device="cuda"), batch)
os.environ["SAP_REAL_FORCE_GENERIC"] = "1"
env = BatchedRealConstellationEnv(2, 12, 20, 3, 3, 4, 3, 0.5, sat_prox_mat=rng.random((2, 12, 20, 3), dtype=np.float32))
batch = batch_for(env, 2)
env.reset(batch)
env.step(th.randint(0, 20, (2, 12), device="cuda"), batch)
th.cuda.synchronize()
print("sanitize case done")
