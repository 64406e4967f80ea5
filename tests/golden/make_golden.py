"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports /root/reference/src through oracle/ref_import.py (gym / orbit-sim / matplotlib stubs),
drives the reference classes on small seeded inputs and stores inputs + outputs as
``tests/golden/*.npz``.  The GPU box has no /root/reference, so tests only read the .npz files.

Every benefit tensor is fp32-representable (generated as float32, handed to the reference as
float64), so the float64 reference and an fp32-input device path see identical numbers.
Tie-heavy cases run the reference under ``stable_argsort`` (numpy's default unstable sort has
a CPU-dependent tie order); tie-free cases run it untouched.
"""
from __future__ import annotations

import contextlib
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import cpu_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402


def _run_real(R, S, M, N, L, lam, T_ctor, actions, stable, task_prios=None):
    """One RealConstellationEnv episode; returns per-step records."""
    n, m, T = S.shape
    ctx = ref_import.stable_argsort() if stable else _null()
    with ctx:
        env = R.real_env.RealConstellationEnv(1, n, m=m, T=T_ctor, N=N, M=M, L=L, lambda_=lam,
                                              sat_prox_mat=S.astype(np.float64), graphs=1, task_prios=task_prios)
        env.reset()
        rec = {"obs": [], "beta": [], "prev": [], "rewards": [], "done": []}

        def snap():
            pre = env.get_pretransition_data()
            rec["obs"].append(np.array(pre["obs"][0], dtype=np.float64))
            rec["beta"].append(np.array(pre["beta"][0], dtype=np.float64))
            rec["prev"].append(np.array(pre["prev_assigns"][0], dtype=np.int64))
            assert np.all(np.array(pre["avail_actions"][0]) == 1)

        snap()
        for t in range(actions.shape[0]):
            r, d, info = env.step(list(actions[t]))
            assert info == {}
            rec["rewards"].append(np.array(r, dtype=np.float64))
            rec["done"].append(bool(d))
            snap()
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(n=env.n, m=env.m, T=env.T, L=env.L, obs_size=env.get_obs_size())
    return out


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def golden_kat1(R):
    """KAT-1: the author's fixture, experiments.py:265-288."""
    S = np.zeros((4, 4, 2), dtype=np.float32)
    S[:, :, 0] = [[5, 0, 0, 1], [2, 0, 0, 0], [3, 1, 4, 2], [1, 3, 0, 10]]
    S[:, :, 1] = 1
    actions = np.array([[0, 0, 2, 3], [1, 1, 1, 3]], dtype=np.int64)
    out = _run_real(R, S, M=2, N=2, L=2, lam=0.5, T_ctor=1, actions=actions, stable=True)
    np.savez_compressed(os.path.join(HERE, "kat1_real.npz"), S=S, actions=actions, M=2, N=2, L_arg=2, T_ctor=1,
             lambda_=0.5, **out)


def golden_real_random(R):
    cases = [
        # name, gen, n, m, T, L, M, N, stable, seed, prios
        ("real_exact_tiefree", "exact", 7, 11, 6, 3, 4, 3, False, 1, False),
        ("real_dense", "dense", 12, 16, 5, 3, 6, 4, False, 2, False),
        ("real_ties", "ties", 9, 14, 6, 3, 4, 3, True, 3, False),
        ("real_reflike", "ref", 10, 24, 8, 3, 6, 3, True, 4, False),
        ("real_prios_L2", "dense", 8, 13, 5, 2, 4, 2, False, 5, True),
        ("real_L1", "exact", 6, 9, 4, 1, 2, 2, True, 6, False),
        ("real_mid", "dense", 24, 40, 4, 3, 10, 10, False, 7, False),
    ]
    for name, gen, n, m, T, L, M, N, stable, seed, prios in cases:
        rng = np.random.default_rng(seed)
        if gen == "exact":
            # tie-free by rejection: distinct row sums are almost surely distinct on the grid plus offsets
            S = O.gen_exact(rng, 1, n, m, T)[0]
            off = (rng.permutation(n * m).reshape(n, m, 1).astype(np.float32) + 1) * np.float32(2.0 ** -20)
            S = (S + off).astype(np.float32)
        elif gen == "dense":
            S = O.gen_dense(rng, 1, n, m, T)[0]
        elif gen == "ties":
            S = O.gen_exact(rng, 1, n, m, T, zero_frac=0.6)[0]
            S = (np.round(S * 8) / 8).astype(np.float32)  # coarse grid -> many duplicates
        else:
            S = O.gen_ref_like(rng, 1, n, m, T)[0]
        tp = (rng.integers(1, 5, size=m).astype(np.float64) * 0.25) if prios else None
        actions = rng.integers(0, m, size=(T, n), dtype=np.int64)
        # make conflicts and "stay on the same task" cases frequent
        actions[:, : n // 2] = actions[:, :1]
        actions[1::2, n // 2:] = actions[0::2, n // 2:][: actions[1::2].shape[0]]
        out = _run_real(R, S, M=M, N=N, L=L, lam=0.5, T_ctor=T, actions=actions, stable=stable, task_prios=tp)
        extra = {} if tp is None else {"task_prios": tp}
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), S=S, actions=actions, M=M, N=N, L_arg=L, T_ctor=T,
                 lambda_=0.5, stable=stable, **extra, **out)


def golden_mock(R):
    for name, n, m, T, L, seed, gen in [("mock_small", 5, 7, 6, 3, 11, "dense"),
                                        ("mock_ref", 10, 10, 12, 3, 12, "ref"),
                                        ("mock_L4", 6, 9, 5, 4, 13, "exact")]:
        rng = np.random.default_rng(seed)
        S = {"dense": O.gen_dense, "ref": O.gen_ref_like, "exact": O.gen_exact}[gen](rng, 1, n, m, T)[0]
        actions = rng.integers(0, m, size=(T, n), dtype=np.int64)
        actions[:, : n // 2] = actions[:, :1]
        np.random.seed(seed)
        env = R.mock_env.MockConstellationEnv(n, m, T, L, 0.5, sat_prox_mat=S.astype(np.float64))
        env.reset()
        prev0 = np.array(env.prev_assigns, dtype=np.int64)
        rec = {"obs": [], "beta": [], "rewards": [], "done": []}

        def snap():
            pre = env.get_pretransition_data()
            assert set(pre) == {"obs", "avail_actions", "beta"}
            rec["obs"].append(np.array(pre["obs"][0], dtype=np.float64))
            rec["beta"].append(np.array(pre["beta"][0], dtype=np.float64))

        snap()
        for t in range(T):
            r, d, info = env.step(list(actions[t]))
            rec["rewards"].append(np.array(r, dtype=np.float64))
            rec["done"].append(bool(d))
            snap()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), S=S, actions=actions, prev0=prev0, L=L, lambda_=0.5,
                 **{k: np.stack(v) for k, v in rec.items()})


def golden_selectors(R):
    import torch as th

    rng = np.random.default_rng(21)
    B, n, A = 3, 5, 9
    args = SimpleNamespace(epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=1000,
                           evaluation_epsilon=0.0, use_mps_action_selection=True, device="cpu",
                           env_args={"M": 4})
    out = {}
    # --- epsilon schedule ---
    sched = R.schedules.DecayThenFlatSchedule(1.0, 0.05, 1000, decay="linear")
    ts = np.array([0, 1, 250, 999, 1000, 5000], dtype=np.int64)
    out["sched_t"] = ts
    out["sched_eps"] = np.array([sched.eval(int(t)) for t in ts])

    # --- classic epsilon greedy with injected draws ---
    q = rng.standard_normal((B, n, A)).astype(np.float32)
    q[0, 0, 3] = q[0, 0, 5] = q[0, 0].max() + 1  # argmax tie -> first index
    avail = rng.random((B, n, A)) > 0.3
    avail[..., 0] |= ~avail.any(-1)
    avail[0, 0, 3] = avail[0, 0, 5] = True
    q[1, 1, 2] = 100.0
    avail[1, 1, 2] = False  # best action masked out
    u_explore = rng.random((B, n), dtype=np.float32)
    u_action = rng.random((B, n), dtype=np.float32)
    sel = R.classic_selectors.EpsilonGreedyActionSelector(args)
    picks = []
    t_envs = [0, 500, 1000]
    u_explore[2, 4] = np.float32(sched.eval(500))  # u == eps exactly -> not random (strict <)
    for t_env in t_envs + ["test"]:
        with _inject(th, R, [u_explore], u_action):
            if t_env == "test":
                a = sel.select_action(th.tensor(q), th.tensor(avail), 0, test_mode=True)
            else:
                a = sel.select_action(th.tensor(q), th.tensor(avail), t_env, test_mode=False)
        picks.append(a.numpy())
    out.update(eg_q=q, eg_avail=avail, eg_u_explore=u_explore, eg_u_action=u_action,
               eg_t_env=np.array(t_envs), eg_actions=np.stack(picks))

    # --- filtered epsilon greedy ---
    M, m, L = 4, 9, 3
    qf = rng.standard_normal((B, n, M + 1)).astype(np.float32)
    qf[0, 1, :M] = -5.0  # baseline wins -> decided by tie noise (or first index)
    qf[0, 2, M] = 0.5  # |base| >= 0.25 -> noise below half ulp -> first non-top index
    qf[0, 2, :M] = 0.0
    beta = O.gen_exact(rng, B, n, m, L)[..., :L]
    beta = (beta + (rng.permutation(B * n * m).reshape(B, n, m, 1) + 1).astype(np.float32) * np.float32(2.0 ** -20))
    beta = beta.astype(np.float32)
    u_tie = rng.random((B, n, m), dtype=np.float32)
    u_explore2 = rng.random((B, n), dtype=np.float32)
    u_action2 = rng.random((B, n), dtype=np.float32)
    fsel = R.filtered_selectors.FilteredEpsilonGreedyActionSelector(args)
    fpicks = []
    for t_env in t_envs + ["test"]:
        with _inject(th, R, [u_tie, u_explore2], u_action2):
            if t_env == "test":
                a = fsel.select_action(th.tensor(qf), th.tensor(avail), 0, test_mode=True, beta=th.tensor(beta))
            else:
                a = fsel.select_action(th.tensor(qf), th.tensor(avail), t_env, test_mode=False, beta=th.tensor(beta))
        fpicks.append(a.numpy())
    out.update(fg_q=qf, fg_beta=beta, fg_u_tie=u_tie, fg_u_explore=u_explore2, fg_u_action=u_action2,
               fg_actions=np.stack(fpicks), fg_M=M)
    out.update(eps_start=1.0, eps_finish=0.05, eps_anneal=1000, eval_eps=0.0)
    np.savez_compressed(os.path.join(HERE, "selectors.npz"), **out)


def golden_sap_selectors(R):
    """The four assignment selectors of the reference (sap_selectors.py, filtered_sap_selectors.py) on seeded inputs.
    th.normal / th.rand_like are replaced so that the reference consumes the recorded draws (one call per env)."""
    import torch as th

    rng = np.random.default_rng(33)
    B, n, m, M, L = 4, 7, 11, 4, 3
    args = SimpleNamespace(epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=1000,
                           evaluation_epsilon=0.0, use_mps_action_selection=False, device="cpu",
                           env_args={"M": M})
    out = dict(sap_B=B, sap_n=n, sap_m=m, sap_M=M)
    q = rng.standard_normal((B, n, m)).astype(np.float32)
    z = rng.standard_normal((B, n, m)).astype(np.float32)
    avail = np.ones((B, n, m), dtype=bool)
    t_envs = [0, 400, 1000]

    class patched:
        def __init__(self, normals, rands):
            self.normals, self.rands = list(normals), list(rands)

        def __enter__(self):
            self._n, self._r = th.normal, th.rand_like
            normals, rands = self.normals, self.rands

            def normal(mean=None, std=None, **k):
                v = normals.pop(0)
                assert tuple(v.shape) == tuple(std.shape)
                return th.tensor(v) * std + mean

            def rand_like(x, *a, **k):
                v = rands.pop(0)
                assert tuple(v.shape) == tuple(x.shape)
                return th.tensor(v, dtype=x.dtype)

            th.normal, th.rand_like = normal, rand_like
            return self

        def __exit__(self, *a):
            th.normal, th.rand_like = self._n, self._r
            return False

    # "sap": Gaussian perturbation + linear_sum_assignment per env
    sel = R.sap_selectors.SequentialAssignmentProblemSelector(args)
    picks = []
    for t_env in t_envs + ["test"]:
        with patched([z[b] for b in range(B)], []):
            a = sel.select_action(th.tensor(q), th.tensor(avail), 0 if t_env == "test" else t_env, test_mode=t_env == "test")
        picks.append(a.numpy().astype(np.int64))
    out.update(sap_q=q, sap_z=z, sap_t_env=np.array(t_envs), sap_actions=np.stack(picks))
    # "epsilon_greedy_sap_test", test branch: assignment of the raw Q-values
    sel2 = R.sap_selectors.EpsilonGreedySAPTestActionSelector(args)
    out["egsap_test_actions"] = sel2.select_action(th.tensor(q), th.tensor(avail), 0, test_mode=True).numpy().astype(np.int64)
    # filtered variants
    qf = rng.standard_normal((B, n, M + 1)).astype(np.float32)
    beta = O.gen_exact(rng, B, n, m, L)[..., :L]
    beta = (beta + (rng.permutation(B * n * m).reshape(B, n, m, 1) + 1).astype(np.float32) * np.float32(2.0 ** -20)).astype(np.float32)
    u_tie = rng.random((B, n, m), dtype=np.float32)
    zf = rng.standard_normal((B, n, m)).astype(np.float32)
    fsel = R.filtered_sap_selectors.FilteredSAPActionSelector(args)
    fpicks = []
    for t_env in t_envs + ["test"]:
        with patched([zf[b] for b in range(B)], [u_tie[b] for b in range(B)]):
            a = fsel.select_action(th.tensor(qf), th.tensor(avail), 0 if t_env == "test" else t_env,
                                   test_mode=t_env == "test", beta=th.tensor(beta))
        fpicks.append(a.numpy().astype(np.int64))
    fsel2 = R.filtered_sap_selectors.FilteredEpsGrSAPTestActionSelector(args)
    with patched([], [u_tie[b] for b in range(B)]):
        a2 = fsel2.select_action(th.tensor(qf), th.tensor(avail), 0, test_mode=True, beta=th.tensor(beta))
    out.update(fsap_q=qf, fsap_beta=beta, fsap_u_tie=u_tie, fsap_z=zf, fsap_actions=np.stack(fpicks),
               fepsgr_test_actions=a2.numpy().astype(np.int64))
    out.update(eps_start=1.0, eps_finish=0.05, eps_anneal=1000, eval_eps=0.0)
    np.savez_compressed(os.path.join(HERE, "sap_selectors.npz"), **out)


def golden_policy_selectors(R):
    """Multinomial / SoftPolicies / FilteredSoftPolicies selectors of the reference.  Categorical.sample is replaced by
    the injected-uniform contract of oracle.sample_categorical and th.rand by recorded draws, so the fixture pins the
    masking, the greedy test branch and the filtered index mapping of the reference code."""
    import torch as th
    from torch.distributions import Categorical

    rng = np.random.default_rng(55)
    B, n, m, M, L = 3, 6, 10, 4, 3
    args = SimpleNamespace(epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=1000, evaluation_epsilon=0.0,
                           use_mps_action_selection=True, device="cpu", env_args={"M": M}, test_greedy=True)
    p = rng.random((B, n, m)).astype(np.float32)
    p[0, 0, :3] = 0.0
    p /= p.sum(-1, keepdims=True)
    avail = rng.random((B, n, m)) > 0.3
    avail[..., 0] |= ~avail.any(-1)
    u = rng.random((B, n), dtype=np.float32)
    u[0, 0], u[0, 1] = 0.0, np.float32(1.0 - 2.0 ** -24)
    orig_sample, orig_rand = Categorical.sample, th.rand
    state = {}

    def sample(dist, sample_shape=th.Size()):
        return th.tensor(O.sample_categorical(dist.probs.numpy(), state["u"]))

    Categorical.sample = sample
    try:
        state["u"] = u
        msel = R.classic_selectors.MultinomialActionSelector(args)
        a_train = msel.select_action(th.tensor(p), th.tensor(avail), 0, test_mode=False).numpy()
        a_test = msel.select_action(th.tensor(p), th.tensor(avail), 0, test_mode=True).numpy()
        ssel = R.classic_selectors.SoftPoliciesSelector(args)
        a_soft = ssel.select_action(th.tensor(p), th.tensor(avail), 0).numpy()
        pf = rng.random((B, n, M + 1)).astype(np.float32)
        pf /= pf.sum(-1, keepdims=True)
        beta = O.gen_exact(rng, B, n, m, L)[..., :L]
        beta = (beta + (rng.permutation(B * n * m).reshape(B, n, m, 1) + 1).astype(np.float32) * np.float32(2.0 ** -20)).astype(np.float32)
        u_rand = rng.random((B, n, m), dtype=np.float32)
        th.rand = lambda *a, **k: th.tensor(u_rand)
        fsel = R.filtered_selectors.FilteredSoftPoliciesSelector(args)
        uf = rng.random((B, n), dtype=np.float32)
        uf[0, :3] = np.float32(0.999)  # force the "anything else" slot
        state["u"] = uf
        a_filt = fsel.select_action(th.tensor(pf), th.tensor(avail), 0, beta=th.tensor(beta)).numpy()
    finally:
        Categorical.sample, th.rand = orig_sample, orig_rand
    np.savez_compressed(os.path.join(HERE, "policy_selectors.npz"), p=p, avail=avail, u=u, multinomial_train=a_train,
                        multinomial_test=a_test, soft=a_soft, pf=pf, beta=beta, u_rand=u_rand, uf=uf, filtered=a_filt, M=M)


def golden_bids(R):
    """bids_as_actions (real_constellation_env.py:140-141, mock_constellation_env.py:121-122): the reference envs stepped
    with a bid matrix per step (scipy turns it into the assignment)."""
    rng = np.random.default_rng(91)
    n, m, T, L, M, N, lam = 6, 9, 4, 3, 4, 3, 0.5
    S = O.gen_dense(rng, 1, n, m, T)[0]
    bids = rng.random((T, n, m)).astype(np.float32)
    out = dict(S=S, bids=bids, L=L, M=M, N=N, lambda_=lam)
    env = R.real_env.RealConstellationEnv(1, n, m=m, T=T, N=N, M=M, L=L, lambda_=lam, sat_prox_mat=S.astype(np.float64),
                                          graphs=1, bids_as_actions=True)
    assert env.scheme["actions"]["vshape"] == (m,)
    env.reset()
    rew, obs = [], [np.array(env.get_obs())]
    for t in range(T):
        r, d, _ = env.step(bids[t].astype(np.float64))
        rew.append(r)
        obs.append(np.array(env.get_obs()))
    out.update(real_rewards=np.array(rew), real_obs=np.stack(obs), real_prev=np.array(env.prev_assigns))
    menv = R.mock_env.MockConstellationEnv(n, m, T, L, lam, bids_as_actions=True, sat_prox_mat=S.astype(np.float64))
    np.random.seed(4)
    menv.reset()
    out["mock_prev0"] = np.array(menv.prev_assigns)
    mrew = []
    for t in range(T):
        r, d, _ = menv.step(bids[t].astype(np.float64))
        mrew.append(r)
    out.update(mock_rewards=np.array(mrew))
    np.savez_compressed(os.path.join(HERE, "bids.npz"), **out)


def golden_haa(R):
    """HAASelector (non_rl_selectors.py:10-50) of the unmodified reference on real-env states: at every step of a short
    episode, the reference's pick from the state fields (beta, prev_assigns) of its own EpisodeBatch."""
    import importlib

    import torch as th

    non_rl = importlib.import_module("action_selectors.non_rl_selectors")
    rng = np.random.default_rng(77)
    n, m, T, L, M, N, lam = 6, 9, 5, 3, 4, 3, 0.5
    S = O.gen_dense(rng, 1, n, m, T)[0]
    env = R.real_env.RealConstellationEnv(1, n, m=m, T=T, N=N, M=M, L=L, lambda_=lam,
                                          sat_prox_mat=S.astype(np.float64), graphs=1)
    args = SimpleNamespace(use_mps_action_selection=False, device="cpu", runner="episode")
    sel = non_rl.HAASelector(args)
    sel.envs = [env]
    batch = R.episode_buffer.EpisodeBatch(env.scheme, {"agents": n}, 1, T + 1, preprocess=env.preprocess, device="cpu")
    env.reset()
    picks, follow = [], rng.integers(0, m, size=(T, n))
    for t in range(T):
        batch.update(env.get_pretransition_data(), ts=t)
        a = sel.select_action(batch[:, t]).numpy().astype(np.int64)[0]
        picks.append(a)
        # alternate between following HAA and a random joint action so that prev_assigns varies
        env.step(a if t % 2 == 0 else follow[t])
    np.savez_compressed(os.path.join(HERE, "haa.npz"), S=S, L=L, M=M, N=N, lambda_=lam, haa_actions=np.stack(picks),
                        follow=follow)


def golden_haal(R):
    """HAALSelector (non_rl_selectors.py:54-145) of the unmodified reference: at every step of a short episode its pick for
    the live env (look-ahead over all time-interval sequences with deep-copied envs and scipy assignments)."""
    import importlib

    import torch as th

    non_rl = importlib.import_module("action_selectors.non_rl_selectors")
    rng = np.random.default_rng(78)
    n, m, T, L, M, N, lam = 6, 9, 7, 3, 4, 3, 0.5
    S = O.gen_dense(rng, 1, n, m, T)[0]
    prios = rng.choice([1.0, 2.0], size=m)
    env = R.real_env.RealConstellationEnv(1, n, m=m, T=T, N=N, M=M, L=L, lambda_=lam, sat_prox_mat=S.astype(np.float64),
                                          graphs=1, task_prios=prios)
    args = SimpleNamespace(use_mps_action_selection=False, device="cpu", runner="episode")
    sel = non_rl.HAALSelector(args)
    sel.envs = [env]
    batch = R.episode_buffer.EpisodeBatch(env.scheme, {"agents": n}, 1, T + 1, preprocess=env.preprocess, device="cpu")
    env.reset()
    picks, prevs, follow = [], [], rng.integers(0, m, size=(T, n))
    for t in range(T):
        batch.update(env.get_pretransition_data(), ts=t)
        prevs.append(np.array(env.prev_assigns).astype(np.int64))
        a = sel.select_action(batch[:, t]).numpy().astype(np.int64)[0]
        picks.append(a)
        env.step(a if t % 2 == 0 else follow[t])
    np.savez_compressed(os.path.join(HERE, "haal.npz"), S=S, L=L, M=M, N=N, lambda_=lam, task_prios=prios,
                        haal_actions=np.stack(picks), prev=np.stack(prevs), follow=follow)


class _inject:
    """Feed injected uniforms to th.rand_like (in call order) and replace Categorical.sample by the
    rank-select contract of oracle.random_available_action."""

    def __init__(self, th, R, rand_like_values, u_action):
        self.th, self.vals, self.u_action = th, list(rand_like_values), u_action

    def __enter__(self):
        from torch.distributions import Categorical

        th = self.th
        self._rl, self._cs, self.Cat = th.rand_like, Categorical.sample, Categorical
        vals, u_action = self.vals, self.u_action

        def rand_like(x, *a, **k):
            v = vals.pop(0)
            assert tuple(v.shape) == tuple(x.shape), (v.shape, x.shape)
            return th.tensor(v, dtype=x.dtype)

        def sample(dist, sample_shape=th.Size()):
            avail = dist.probs.numpy() > 0
            return th.tensor(O.random_available_action(avail, u_action))

        th.rand_like = rand_like
        Categorical.sample = sample
        return self

    def __exit__(self, *a):
        self.th.rand_like = self._rl
        self.Cat.sample = self._cs
        return False


def golden_buffer(R):
    """EpisodeBatch.update casting + OneHot + ReplayBuffer ring insert (episode_buffer.py)."""
    import torch as th

    rng = np.random.default_rng(31)
    n, m, L, T = 3, 5, 2, 4
    env_scheme = {
        "obs": {"vshape": 6, "group": "agents", "dtype": th.float16},
        "actions": {"vshape": (1,), "group": "agents", "dtype": th.int16},
        "avail_actions": {"vshape": (m,), "group": "agents", "dtype": th.bool},
        "rewards": {"vshape": (n,), "dtype": th.float16},
        "terminated": {"vshape": (1,), "dtype": th.bool},
        "prev_assigns": {"vshape": (n,), "dtype": th.int16, "part_of_state": True},
        "beta": {"vshape": (n, m, L), "dtype": th.float16, "part_of_state": True},
    }
    groups = {"agents": n}
    pre = {"actions": ("actions_onehot", [R.transforms.OneHot(out_dim=m)])}
    EB, RB = R.episode_buffer.EpisodeBatch, R.episode_buffer.ReplayBuffer
    rb = RB(dict(env_scheme), groups, 5, T + 1, preprocess=pre, device="cpu")
    eps = []
    raw = []
    for e in range(4):
        B = 2
        batch = EB(dict(env_scheme), groups, B, T + 1, preprocess=pre, device="cpu")
        rec = {"obs": rng.standard_normal((B, T + 1, n, 6)) * 3,
               "rewards": rng.standard_normal((B, T, n)) * 2,
               "actions": rng.integers(0, m, size=(B, T, n)),
               "beta": rng.random((B, T + 1, n, m, L))}
        for t in range(T + 1):
            batch.update({"obs": [rec["obs"][b, t] for b in range(B)],
                          "beta": [rec["beta"][b, t] for b in range(B)],
                          "avail_actions": [[[1] * m] * n for b in range(B)],
                          "prev_assigns": [np.arange(n) for b in range(B)]}, ts=t)
            if t < T:
                batch.update({"actions": th.tensor(rec["actions"][:, t]),
                              "rewards": [(list(rec["rewards"][b, t]),) for b in range(B)],
                              "terminated": [(t == T - 1,) for b in range(B)]}, ts=t)
        rb.insert_episode_batch(batch)
        raw.append(rec)
        eps.append({k: v.numpy().copy() if v.dtype != th.bool else v.numpy().copy()
                    for k, v in batch.data.transition_data.items()})
    out = {}
    for e, (rec, ep) in enumerate(zip(raw, eps)):
        for k, v in rec.items():
            out[f"raw{e}_{k}"] = v
        for k, v in ep.items():
            out[f"ep{e}_{k}"] = v
    for k, v in rb.data.transition_data.items():
        out[f"rb_{k}"] = v.numpy().copy()
    out["rb_buffer_index"] = rb.buffer_index
    out["rb_episodes_in_buffer"] = rb.episodes_in_buffer
    out["max_t_filled"] = int(rb.max_t_filled())
    np.savez_compressed(os.path.join(HERE, "buffer.npz"), n=n, m=m, L=L, T=T, **out)


def main():
    assert ref_import.reference_available(), "needs /root/reference"
    R = ref_import.ref_modules()
    golden_kat1(R)
    golden_real_random(R)
    golden_mock(R)
    golden_selectors(R)
    golden_sap_selectors(R)
    golden_haa(R)
    golden_haal(R)
    golden_bids(R)
    golden_policy_selectors(R)
    golden_buffer(R)
    golden_runner(R)
    golden_parallel_runner(R)
    golden_power_envs(R)
    golden_proximities(R)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))




def golden_runner(R):
    """The reference's EpisodeRunner + BasicMAC + RNNAgent + epsilon_greedy end to end (runners/episode_runner.py:60-127),
    with injected selector draws.  Seeds are retried until every greedy decision has a top-2 Q gap > 1e-3, so that
    fp32 GEMM rounding differences between the reference's CPU agent and a GPU agent cannot flip an argmax."""
    import torch as th
    from types import SimpleNamespace as SN

    class Log:
        def __init__(self):
            self.stats = {}

        def log_stat(self, k, v, t):
            self.stats.setdefault(k, []).append((t, float(v)))

    for name, env_name, n, m, T, extra in [("runner_mock", "mock_constellation_env", 10, 10, 20, {}),
                                           ("runner_real", "real_constellation_env", 8, 12, 10, dict(M=4, N=3))]:
        for seed in range(100):
            rng = np.random.default_rng(1000 + seed)
            S = O.gen_dense(rng, 1, n, m, T)[0]
            if env_name == "mock_constellation_env":
                env_args = dict(n=n, m=m, T=T, L=3, lambda_=0.5, sat_prox_mat=S.astype(np.float64), seed=0)
            else:
                env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, L=3, lambda_=0.5, N=extra["N"], M=extra["M"],
                                sat_prox_mat=S.astype(np.float64), graphs=1, seed=0)
            args = SN(env=env_name, env_args=env_args, batch_size_run=1, use_mps_action_selection=True, device="cpu",
                      mac="basic_mac", render=False, test_nepisode=1, runner_log_interval=1, agent="rnn", hidden_dim=64,
                      use_rnn=False, agent_output_type="q", action_selector="epsilon_greedy", epsilon_start=0.4,
                      epsilon_finish=0.4, epsilon_anneal_time=1, evaluation_epsilon=0.0, obs_agent_id=True,
                      obs_last_action=True, n=n, m=m, T=T)
            log = Log()
            runner = R.episode_runner.EpisodeRunner(args, log)
            env = runner.get_env()
            groups = {"agents": n}
            buffer = R.episode_buffer.ReplayBuffer(env.scheme, groups, 2, T + 1, preprocess=env.preprocess, device="cpu")
            th.manual_seed(seed)
            mac = R.basic_controller.BasicMAC(buffer.scheme, groups, args)
            runner.setup(scheme=env.scheme, groups=groups, preprocess=env.preprocess, mac=mac)
            u_explore = rng.random((T, 1, n), dtype=np.float32)
            u_action = rng.random((T, 1, n), dtype=np.float32)
            gaps = []
            sel = mac.action_selector
            orig = sel.select_action
            step = [0]

            def wrapped(agent_inputs, avail_actions, t_env, test_mode=False, beta=None):
                top2 = th.topk(agent_inputs, 2, dim=-1).values
                gaps.append(float((top2[..., 0] - top2[..., 1]).min()))
                with _inject(th, R, [u_explore[step[0]]], u_action[step[0]]):
                    a = orig(agent_inputs, avail_actions, t_env, test_mode=test_mode, beta=beta)
                step[0] += 1
                return a

            sel.select_action = wrapped
            np.random.seed(seed)
            with th.no_grad():
                batch = runner.run(test_mode=False)
            if min(gaps) > 1e-3:
                break
        else:
            raise RuntimeError("no seed with a safe Q gap")
        out = {f"td_{k}": v.numpy().copy() for k, v in batch.data.transition_data.items()}
        out.update({f"w_{k}": v.numpy().copy() for k, v in mac.agent.state_dict().items()})
        prev0 = out["td_obs"][0, 0] * 0  # placeholder keeps key order stable
        extra_out = {}
        if env_name == "mock_constellation_env":
            np.random.seed(seed)
            extra_out["prev0"] = np.random.choice(m, n, replace=False)  # what env.reset drew (mock_constellation_env.py:105)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), S=S, n=n, m=m, T=T, L=3, lambda_=0.5, M=extra.get("M", 0),
                            N=extra.get("N", 0), u_explore=u_explore, u_action=u_action, eps=0.4, min_gap=min(gaps),
                            t_env_after=runner.t_env, return_mean=log.stats["return_mean"][-1][1], **extra_out, **out)


def _drive_power_like(env, T, actions, stable):
    """reset + T steps of a power-type reference env; returns what the buffer would receive, plus the reset draw."""
    with (ref_import.stable_argsort() if stable else contextlib.nullcontext()):
        env.reset()
        prev0 = np.array(env.prev_assigns).copy()
        rec = {k: [] for k in ("obs", "beta", "prev", "power", "rewards", "done")}

        def snap():
            pre = env.get_pretransition_data()
            rec["obs"].append(np.array(pre["obs"][0], dtype=np.float64))
            rec["beta"].append(np.array(pre["beta"][0], dtype=np.float64))
            rec["prev"].append(np.array(pre["prev_assigns"][0]).astype(np.int64))
            rec["power"].append(np.array(pre["power_states"][0], dtype=np.float64))

        snap()
        for t in range(T):
            r, d, _ = env.step(actions[t])
            rec["rewards"].append(np.array(r, dtype=np.float64))
            rec["done"].append(bool(d))
            snap()
    out = {k: np.array(v) for k, v in rec.items()}
    out["prev0"] = prev0
    return out


def golden_power_envs(R):
    """RealPowerConstellationEnv (envs/real_power_constellation_env.py) and InterferenceConstellationEnv
    (envs/interference_constellation_env.py), unmodified.  Benefits carry inactive tasks (zeros) so that agents recharge
    as well as drain; the episode is long enough for agents to run out of power through the 5.55e-17 residue
    (SURVEY.md Q9).  The interference env builds itself from the orbit simulator only, so the simulator class in its
    module namespace is replaced by a stand-in that hands out a given proximity tensor and neighbour matrix."""
    import importlib

    power_mod = importlib.import_module("envs.real_power_constellation_env")
    inter_mod = importlib.import_module("envs.interference_constellation_env")
    rng = np.random.default_rng(77)
    n, m, T, L, M, N, lam = 8, 12, 14, 3, 4, 3, 0.5
    S = O.gen_dense(rng, 1, n, m, T)[0]
    S[:, ::4] = 0.0                      # tasks nobody can see: choosing them recharges
    S[2] *= (rng.random(m) < 0.5)[:, None]  # an agent with many invisible tasks
    prios = rng.choice([1.0, 1.0, 1.0, 5.0], size=m)
    acts = rng.integers(0, m, size=(T, n))
    acts[:, :3] = acts[:, :1]            # conflicts
    acts[:, 5] = 1                       # an agent that always works a visible task: dies at t = 5, goes negative
    np.random.seed(11)
    env = power_mod.RealPowerConstellationEnv(1, n, m, T, N, M, L, lam, sat_prox_mat=S.astype(np.float64), graphs=1,
                                              task_prios=prios)
    out = _drive_power_like(env, T, acts, stable=True)   # zero-benefit tasks tie: stable rule
    np.savez_compressed(os.path.join(HERE, "power_env.npz"), S=S, n=n, m=m, T=T, L=L, M=M, N=N, lambda_=lam, task_prios=prios,
                        actions=acts, obs_size=env.get_obs_size(), **out)

    # ---- interference env
    neighbor = (rng.random((m, m)) < 0.3).astype(np.float64)
    neighbor = np.maximum(neighbor, neighbor.T)
    np.fill_diagonal(neighbor, 1.0)      # a region neighbours itself (the "- 1" of :331 removes the self-conflict)
    bands = rng.integers(0, 3, size=n)

    class _Sim:
        def __init__(self, num_planes, num_sats_per_plane, T=None):
            self.neighbor_matrix, self.graphs = neighbor, [1]

        def get_proximities_for_coverage_tasks(self, res):
            return S.astype(np.float64)

    saved = inter_mod.HighPerformanceConstellationSim
    inter_mod.HighPerformanceConstellationSim = _Sim
    try:
        np.random.seed(12)
        ienv = inter_mod.InterferenceConstellationEnv(1, n, 2, T, N, M, L, lam, task_prios=prios, sat_freq_bands=bands)
        iout = _drive_power_like(ienv, T, acts, stable=True)
    finally:
        inter_mod.HighPerformanceConstellationSim = saved
    np.savez_compressed(os.path.join(HERE, "interference_env.npz"), S=S, n=n, m=m, T=T, L=L, M=M, N=N, lambda_=lam,
                        task_prios=prios, actions=acts, neighbor_matrix=neighbor, sat_freq_bands=bands,
                        obs_size=ienv.get_obs_size(), **iout)


def golden_proximities(R=None):
    """calc_fov_based_proximities_fast (envs/HighPerformanceConstellationSim.py:308-327), the reference's own function: the
    module needs poliastro at import time, so the function's source is cut out of the unmodified file with ``ast`` and
    executed on its own (it is pure numpy).  Satellites on circular 550 km tracks, some passing over tasks."""
    import ast

    with open(os.path.join(ref_import.REFERENCE_SRC, "envs", "HighPerformanceConstellationSim.py")) as fh:
        tree = ast.parse(fh.read())
    fn = [x for x in tree.body if isinstance(x, ast.FunctionDef) and x.name == "calc_fov_based_proximities_fast"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "HighPerformanceConstellationSim.py", "exec"), ns)
    f = ns["calc_fov_based_proximities_fast"]
    rng = np.random.default_rng(5)
    n, m, T = 12, 20, 9

    def unit(k):
        v = rng.standard_normal((k, 3))
        return v / np.linalg.norm(v, axis=1, keepdims=True)

    task = unit(m) * 6371.0
    base = unit(n)
    base[:6] = task[:6] / 6371.0 + 0.05 * rng.standard_normal((6, 3))
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    axis = unit(n)
    sat = np.zeros((n, 3, T))
    for i in range(n):
        for k in range(T):
            ang = 0.01 * k
            v = base[i] * np.cos(ang) + np.cross(axis[i], base[i]) * np.sin(ang) + axis[i] * np.dot(axis[i], base[i]) * (1 - np.cos(ang))
            sat[i, :, k] = v / np.linalg.norm(v) * 6921.0
    fov = 60.0
    sig2 = np.sqrt(-(fov ** 2) / (2 * np.log(0.05))) ** 2          # :100-101
    prox = np.zeros((n, m, T))
    for i in range(n):
        for j in range(m):
            for k in range(T):
                prox[i, j, k] = f(sat[i, :, k], task[j], fov, sig2)
    assert (prox > 0).mean() > 0.02
    np.savez_compressed(os.path.join(HERE, "proximities.npz"), sat_r=sat, task_r=task, fov=fov, sigma_2=sig2, prox=prox)


def golden_parallel_runner(R):
    """The reference's ParallelRunner (runners/parallel_runner.py:12-243: one forked env process per env, pickled pre-/post-
    transition data over Pipes) + BasicMAC + RNNAgent + epsilon_greedy with B = 4 envs and injected selector draws.  Keeps
    what that runner really writes, quirks included: `terminated` comes from the truthiness of the reward
    list collected so far (:181-187: False for the first env of the batch, True for every other env, at every step), and one
    extra action selection is made and stored at t = T (:131-139) before the loop notices that every env has finished."""
    import torch as th
    from types import SimpleNamespace as SN

    class Log:
        def __init__(self):
            self.stats = {}

        def log_stat(self, k, v, t):
            self.stats.setdefault(k, []).append((t, float(v)))

    import contextlib
    import io

    B, n, m, T, M, N, L = 4, 8, 12, 10, 4, 3, 3
    for seed in range(200):
        rng = np.random.default_rng(5000 + seed)
        S = O.gen_dense(rng, 1, n, m, T)[0]
        env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, L=L, lambda_=0.5, N=N, M=M,
                        sat_prox_mat=S.astype(np.float64), graphs=1, seed=0)
        args = SN(env="real_constellation_env", env_args=env_args, batch_size_run=B, use_mps_action_selection=True, device="cpu",
                  mac="basic_mac", render=False, test_nepisode=B, runner_log_interval=1, agent="rnn", hidden_dim=64,
                  use_rnn=False, agent_output_type="q", action_selector="epsilon_greedy", epsilon_start=0.4,
                  epsilon_finish=0.4, epsilon_anneal_time=1, evaluation_epsilon=0.0, obs_agent_id=True,
                  obs_last_action=True, n=n, m=m, T=T)
        log = Log()
        with contextlib.redirect_stdout(io.StringIO()):
            runner = R.parallel_runner.ParallelRunner(args, log)
            env = runner.get_env()
        groups = {"agents": n}
        buffer = R.episode_buffer.ReplayBuffer(env.scheme, groups, 2 * B, T + 1, preprocess=env.preprocess, device="cpu")
        th.manual_seed(seed)
        mac = R.basic_controller.BasicMAC(buffer.scheme, groups, args)
        with contextlib.redirect_stdout(io.StringIO()):
            runner.setup(scheme=env.scheme, groups=groups, preprocess=env.preprocess, mac=mac)
        u_explore = rng.random((T + 1, B, n), dtype=np.float32)   # T + 1 selections: the extra one at t = T
        u_action = rng.random((T + 1, B, n), dtype=np.float32)
        gaps, step = [], [0]
        sel = mac.action_selector
        orig = sel.select_action

        def wrapped(agent_inputs, avail_actions, t_env, test_mode=False, beta=None):
            top2 = th.topk(agent_inputs, 2, dim=-1).values
            gaps.append(float((top2[..., 0] - top2[..., 1]).min()))
            with _inject(th, R, [u_explore[step[0]]], u_action[step[0]]):
                a = orig(agent_inputs, avail_actions, t_env, test_mode=test_mode, beta=beta)
            step[0] += 1
            return a

        sel.select_action = wrapped
        with th.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            batch = runner.run(test_mode=False)
            runner.close_env()
        assert step[0] == T + 1
        if min(gaps) > 1e-3:
            break
    else:
        raise RuntimeError("no seed with a safe Q gap")
    out = {f"td_{k}": v.numpy().copy() for k, v in batch.data.transition_data.items()}
    out.update({f"w_{k}": v.numpy().copy() for k, v in mac.agent.state_dict().items()})
    np.savez_compressed(os.path.join(HERE, "runner_parallel.npz"), S=S, B=B, n=n, m=m, T=T, L=L, lambda_=0.5, M=M, N=N,
                        u_explore=u_explore, u_action=u_action, eps=0.4, min_gap=min(gaps), t_env_after=runner.t_env,
                        return_mean=log.stats["return_mean"][-1][1], return_std=log.stats["return_std"][-1][1], **out)


if __name__ == "__main__":
    if "--runner-only" in sys.argv:
        golden_runner(ref_import.ref_modules())
    elif "--prox-only" in sys.argv:
        ref_import.install()
        golden_proximities()
    elif "--haal-only" in sys.argv:
        golden_haal(ref_import.ref_modules())
    elif "--power-only" in sys.argv:
        golden_power_envs(ref_import.ref_modules())
    elif "--parallel-only" in sys.argv:
        golden_parallel_runner(ref_import.ref_modules())
    else:
        main()
