"""Race / determinism stress for the bench-shape env kernel (compute-sanitizer is not available on the GPU pool):
several full waves of environments, several benefit laws, the second-generation kernel run twice and compared byte for
byte with itself and with the first-generation kernel (different shared-memory layout, same contract), through the fused
step and through obs-ahead + selection-in-step.   python profiles/stress_paths.py [rounds]"""
import sys

import torch as th

from marl_sap_b200 import _lib
from marl_sap_b200.components.episode_buffer import EpisodeBatch
from marl_sap_b200.envs.batched import BatchedRealConstellationEnv, real_obs_size, real_scheme


def laws(B, n, m, T, g):
    dense = th.rand(B, n, m, T, generator=g)
    ties = (th.rand(B, n, m, T, generator=g) * 4).round() / 4
    sparse = th.rand(B, n, m, T, generator=g) * (th.rand(B, n, m, T, generator=g) < 0.25)
    tiny = th.rand(B, n, m, T, generator=g) * 1e-6 + 0.5          # near-ties: sums differ in the last bits
    neg = th.rand(B, n, m, T, generator=g) - 0.3
    return {"dense": dense, "ties": ties, "sparse": sparse, "near_ties": tiny, "negative": neg}


def rollout(path, S, acts, ahead):
    B, n, m, T = S.shape
    with _lib.select_real_kernel(path):
        env = BatchedRealConstellationEnv(B, n, m, T, 3, 10, 10, 0.5, sat_prox_mat=S)
        scheme, pre = real_scheme(n, m, 3, real_obs_size(10, 10, 3))
        batch = EpisodeBatch(scheme, {"agents": n}, B, T + 1, preprocess=pre, device="cuda",
                             lazy=("beta", "avail_actions", "actions_onehot"))
        ain = [th.zeros(B, n, env.obs_size, device="cuda") for _ in range(2)]
        batch.agent_in = ain[0]
        env.reset(batch)
        snaps = [ain[0].clone()]
        for t in range(T):
            nxt = ain[(t + 1) % 2]
            if ahead:
                env.obs_ahead(batch, agent_in=nxt)
            env.step(acts[t], batch, agent_in=nxt)
            snaps.append(nxt.clone())
        th.cuda.synchronize()
        td = batch.data.transition_data
        return (td["obs"].clone(), td["rewards"].clone(), td["prev_assigns"].clone(), th.stack(snaps), env.ep_return.clone(),
                env.top.clone())


def main(rounds):
    B, n, m, T = 1400, 100, 100, 6
    bad = 0
    for r in range(rounds):
        g = th.Generator().manual_seed(100 + r)
        for name, S in laws(B, n, m, T, g).items():
            S = S.cuda()
            acts = [th.randint(0, m, (B, n), generator=g).cuda() for _ in range(T)]
            ref = rollout(_lib.REAL_PATH_FAST_GEN1, S, acts, False)
            # AUTO = the instantiation with 100 x 100 compiled in; FAST_RUNTIME_SHAPE = the same kernel reading n, m at run time
            for path, ahead in ((_lib.REAL_PATH_AUTO, False), (_lib.REAL_PATH_AUTO, True), (_lib.REAL_PATH_AUTO, False),
                                (_lib.REAL_PATH_AUTO, True), (_lib.REAL_PATH_FAST_RUNTIME_SHAPE, False),
                                (_lib.REAL_PATH_FAST_RUNTIME_SHAPE, True)):
                got = rollout(path, S, acts, ahead)
                ok = all(th.equal(a, b) for a, b in zip(ref, got))
                bad += not ok
                print(f"round {r} {name:10s} path={path} ahead={ahead}: {'ok' if ok else 'MISMATCH'}", flush=True)
    print("stress:", "PASS" if bad == 0 else f"{bad} FAILURES")
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 2) else 0)
