"""Per-CTA phase timeline of the bench env kernel (needs a -DSAP_ABLATE build: SAP_ABLATE=1 python -m marl_sap_b200._build).
Thread 0 of every CTA records %globaltimer at the phase boundaries of one launch; this script reports the phase durations per
CTA and how many CTAs are in the DRAM-heavy phases at the same time (are the CTAs of a wave in lockstep?).
    SAP_DEBUG_SKIP_REDO=99 python profiles/phase_timeline.py [B]"""
import sys

import numpy as np
import torch as th

from marl_sap_b200.components.episode_buffer import EpisodeBatch
from marl_sap_b200.envs.batched import BatchedRealConstellationEnv, real_obs_size, real_scheme

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n = m = 100
T = 8
g = th.Generator().manual_seed(0)
S = th.rand(B, n, m, T, generator=g).cuda()
env = BatchedRealConstellationEnv(B, n, m, T, 3, 10, 10, 0.5, sat_prox_mat=S)
scheme, pre = real_scheme(n, m, 3, real_obs_size(10, 10, 3))
batch = EpisodeBatch(scheme, {"agents": n}, B, T + 1, preprocess=pre, device="cuda", lazy=("beta", "avail_actions", "actions_onehot"))
batch.agent_in = th.zeros(B, n, env.obs_size, device="cuda")
env.scratch = th.zeros(B * 16, dtype=th.float64, device="cuda")
env.reset(batch)
for t in range(4):
    env.step(th.randint(0, m, (B, n), generator=g).cuda(), batch)
th.cuda.synchronize()
ts = env.scratch.view(th.int64).view(B, 16).cpu().numpy()
if ts[:, 7].max() == 0:
    sys.exit("no timestamps: build with SAP_ABLATE=1 and run with SAP_DEBUG_SKIP_REDO=99")
t0 = ts[:, 0].min()
rel = (ts[:, :8] - t0) / 1e3  # us
names = ["reward", "key pass", "task lists", "rivals", "other", "refill", "gather+store"]
dur = np.diff(rel, axis=1)
print(f"B = {B}: kernel span {rel[:, 7].max():.1f} us; per-CTA total: mean {np.mean(rel[:, 7] - rel[:, 0]):.1f} us, "
      f"p5 {np.percentile(rel[:, 7] - rel[:, 0], 5):.1f}, p95 {np.percentile(rel[:, 7] - rel[:, 0], 95):.1f}")
for k, nm in enumerate(names):
    d = dur[:, k]
    print(f"  {nm:13s} mean {d.mean():6.2f} us  p5 {np.percentile(d, 5):6.2f}  p50 {np.percentile(d, 50):6.2f}  p95 {np.percentile(d, 95):6.2f}")
cyc = ts[:, 9:11].astype(np.float64)
print(f"  inside the last phase: gather {cyc[:, 0].mean() / 1965:.2f} us, stores {cyc[:, 1].mean() / 1965:.2f} us (SM cycles / 1965 MHz)")
# concurrency: at 200 sample times, how many CTAs are in each phase
span = rel[:, 7].max()
grid = np.linspace(0.1 * span, 0.9 * span, 200)
occ = np.zeros((len(grid), len(names)))
for k in range(len(names)):
    lo, hi = rel[:, k], rel[:, k + 1]
    occ[:, k] = ((lo[None, :] <= grid[:, None]) & (grid[:, None] < hi[None, :])).sum(1)
tot = occ.sum(1)
print("  CTAs resident (mean over the middle 80 % of the kernel):", round(tot.mean(), 1))
for k, nm in enumerate(names):
    print(f"  in {nm:13s}: mean {occ[:, k].mean():6.1f}  min {occ[:, k].min():5.0f}  max {occ[:, k].max():5.0f}  (share of time of a CTA: {dur[:, k].mean() / dur.sum(1).mean():.2f})")
# start-time clustering: spread of CTA start times within each group of 444 consecutive starts
order = np.sort(rel[:, 0])
waves = [order[i:i + 444] for i in range(0, len(order), 444)]
print("  start-time spread per 444 CTAs (us):", [round(float(w.max() - w.min()), 1) for w in waves][:12])
