"""CPU oracle: a numpy restatement of marl_sap's rollout hot path.

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module.  The
product package ``marl_sap_b200`` never does; it fails loudly when its CUDA library is
missing.

Parity pin: the reference repo ships no tests.  This oracle is pinned
against the UNMODIFIED reference imported through ``oracle/ref_import.py`` by
``tests/golden/make_golden.py``; the recorded vectors live in ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` replays them.  KAT-1 is the author's own fixture
(experiments.py:265-288).

All arithmetic is float64, like the reference.  Everything is vectorised over a leading
env dimension ``B`` (the reference handles one env per object).  Ties in every top-k
follow the canonical *stable* rule:
    descending pick  == np.argsort(-x, kind="stable")[:k]   -> (value desc, index asc)
    ascending  pick  == np.argsort(x,  kind="stable")[-k:]  -> (value asc,  index asc), last k

Layouts (reference layouts):  S = sat_prox_mat [B, n, m, T];  beta [B, n, m, L] (real) or
[B, n, m] (mock);  obs [B, n, obs_size];  actions / prev_assigns [B, n] int64.
"""
from __future__ import annotations

import numpy as np

NEG_INF = -np.inf


# ------------------------------------------------------------------ epsilon schedule
def epsilon_linear(start: float, finish: float, anneal_time: float, t_env: float) -> float:
    """DecayThenFlatSchedule.eval, decay="linear" (components/epsilon_schedules.py:12-23)."""
    delta = (start - finish) / anneal_time
    return max(finish, start - delta * t_env)


# ------------------------------------------------------------------ shared reward rule
def _counts(actions: np.ndarray, m: int) -> np.ndarray:
    """Histogram of chosen tasks per env (real_constellation_env.py:145-147)."""
    B, n = actions.shape
    cnt = np.zeros((B, m), dtype=np.int64)
    np.add.at(cnt, (np.repeat(np.arange(B), n), actions.reshape(-1)), 1)
    return cnt


def _default_T_trans(m: int) -> np.ndarray:
    return np.ones((m, m)) - np.eye(m)


def _split_rewards(bh: np.ndarray, actions: np.ndarray, cnt: np.ndarray) -> np.ndarray:
    """r_i = bh_i / cnt[a_i] if bh_i > 0 else bh_i  (real_constellation_env.py:154-160)."""
    c = np.take_along_axis(cnt, actions, axis=1).astype(np.float64)
    return np.where(bh > 0, bh / c, bh)


# ------------------------------------------------------------------ RealConstellationEnv
def real_window(S: np.ndarray, k: int, L: int, task_prios: np.ndarray | None = None) -> np.ndarray:
    """beta_k[b,i,j,l] = S[b,i,j,k+l] * P[j], zero past T (real_constellation_env.py:127,167-170)."""
    B, n, m, T = S.shape
    beta = np.zeros((B, n, m, L), dtype=np.float64)
    eff = max(0, min(L, T - k))
    if eff > 0:
        beta[..., :eff] = S[..., k:k + eff]
    if task_prios is not None:
        beta = beta * np.asarray(task_prios, dtype=np.float64)[None, None, :, None]
    return beta


def real_obs_size(M: int, N: int, L: int) -> int:
    """real_constellation_env.py:263-265."""
    return M * L + N * M * L + N * M // 2 * L + M


def real_beta_hat_chosen(beta, prev, actions, lambda_, T_trans=None):
    """beta_hat[i, a_i, 0] only (real_constellation_env.py:282-328 evaluated at the chosen entry)."""
    B, n, m, L = beta.shape
    if T_trans is None:
        pen = (actions != prev).astype(np.float64)
    else:
        pen = np.asarray(T_trans, dtype=np.float64)[prev, actions]
    idx = actions[..., None, None]
    chosen = np.take_along_axis(beta, np.broadcast_to(idx, (B, n, 1, L)), axis=2)[:, :, 0, :]  # [B,n,L]
    meaningful = (chosen.sum(-1) > 1e-12).astype(np.float64)
    return chosen[..., 0] - lambda_ * pen * meaningful


def real_beta_hat_full(beta, prev, lambda_, T_trans=None):
    """Full beta_hat tensor (real_constellation_env.py:282-328), vectorised over B."""
    B, n, m, L = beta.shape
    Tt = _default_T_trans(m) if T_trans is None else np.asarray(T_trans, dtype=np.float64)
    pen = Tt[prev]  # [B,n,m]  == onehot(prev) @ T_trans
    meaningful = beta.sum(-1) > 1e-12
    out = beta.copy()
    out[..., 0] = out[..., 0] - lambda_ * pen * meaningful
    return out


def real_build_obs(beta: np.ndarray, prev: np.ndarray, M: int, N: int, return_indices: bool = False,
                   chunk: int = 64):
    """_build_obs for not-done envs (real_constellation_env.py:177-225), stable ties.

    Returns obs [B, n, obs_size] float64 (and the index sets when asked).
    """
    assert M % 2 == 0, "reference's -M//2 slice is only consistent with get_obs_size for even M"
    B, n, m, L = beta.shape
    assert n > N and m >= M + M // 2
    H = M // 2
    osz = real_obs_size(M, N, L)
    obs = np.empty((B, n, osz), dtype=np.float64)
    tops = np.empty((B, n, M), dtype=np.int64)
    nbrs = np.empty((B, n, N), dtype=np.int64)
    others = np.empty((B, n, N, H), dtype=np.int64)
    ar_n = np.arange(n)
    for b0 in range(0, B, chunk):
        bt = beta[b0:b0 + chunk]
        pv = prev[b0:b0 + chunk]
        Bc = bt.shape[0]
        tot = bt.sum(-1)                                                   # :190
        top = np.argsort(-tot, axis=-1, kind="stable")[..., :M]           # :198  [Bc,n,M]
        # score[b,i,a] = max_q tot[b,a,top[b,i,q]]                        # :203-204
        bi = np.arange(Bc)[:, None, None, None]
        sc = tot[bi, ar_n[None, None, :, None], top[:, :, None, :]].max(-1)  # [Bc,n(i),n(a)]
        sc[:, ar_n, ar_n] = NEG_INF                                        # :205
        nbr = np.argsort(-sc, axis=-1, kind="stable")[..., :N]            # :206  [Bc,n,N]
        # rivals' rows with agent i's top-M tasks masked out               # :212-217
        nb_tot = tot[np.arange(Bc)[:, None, None], nbr]                    # [Bc,n,N,m] (copy)
        np.put_along_axis(nb_tot, np.broadcast_to(top[:, :, None, :], (Bc, n, N, M)), NEG_INF, axis=-1)
        other = np.argsort(nb_tot, axis=-1, kind="stable")[..., m - H:]   # [Bc,n,N,H]
        # gathers                                                          # :199,209,219
        b3 = np.arange(Bc)[:, None, None]
        local = bt[b3, ar_n[None, :, None], top]                           # [Bc,n,M,L]
        b4 = np.arange(Bc)[:, None, None, None]
        neigh = bt[b4, nbr[:, :, :, None], top[:, :, None, :]]             # [Bc,n,N,M,L]
        neigh_other = bt[b4, nbr[:, :, :, None], other]                    # [Bc,n,N,H,L]
        flags = (top == pv[:, :, None]).astype(np.float64)                 # :222
        obs[b0:b0 + Bc] = np.concatenate(
            [local.reshape(Bc, n, -1), neigh.reshape(Bc, n, -1), neigh_other.reshape(Bc, n, -1), flags], axis=-1)
        tops[b0:b0 + Bc], nbrs[b0:b0 + Bc], others[b0:b0 + Bc] = top, nbr, other
    if return_indices:
        return obs, tops, nbrs, others
    return obs


class RealState:
    """Mutable per-batch state of B RealConstellationEnv instances."""

    def __init__(self, S, L, M, N, lambda_, task_prios=None, T_trans=None, T_ctor=None):
        self.S = np.asarray(S, dtype=np.float64)
        self.B, self.n, self.m, self.T = self.S.shape
        # L = min(L, T_ctor) is computed BEFORE T is overridden by S.shape[2] (:38 vs :58-60)
        self.L = min(L, T_ctor if T_ctor is not None else self.T)
        self.M, self.N, self.lambda_ = M, N, lambda_
        self.task_prios = None if task_prios is None else np.asarray(task_prios, dtype=np.float64)
        self.T_trans = None if T_trans is None else np.asarray(T_trans, dtype=np.float64)
        self.obs_size = real_obs_size(M, N, self.L)
        self.k = 0
        self.done = False
        self.beta = None
        self.prev = None
        self.obs = None

    def reset(self):
        """real_constellation_env.py:116-133."""
        self.k, self.done = 0, False
        self.beta = real_window(self.S, 0, self.L, self.task_prios)
        self.prev = np.broadcast_to(np.arange(self.n), (self.B, self.n)).copy()
        self.obs = real_build_obs(self.beta, self.prev, self.M, self.N)
        return self.obs

    def step(self, actions):
        """real_constellation_env.py:135-175. Returns rewards [B,n] float64, done bool."""
        a = np.asarray(actions, dtype=np.int64)
        cnt = _counts(a, self.m)
        bh = real_beta_hat_chosen(self.beta, self.prev, a, self.lambda_, self.T_trans)
        rewards = _split_rewards(bh, a, cnt)
        self.k += 1
        self.done = self.k >= self.T
        self.prev = a.copy()
        if self.done:                                                       # :226-228
            self.beta = np.zeros((self.B, self.n, self.m, self.L))
            self.obs = np.zeros((self.B, self.n, self.obs_size))
        else:
            self.beta = real_window(self.S, self.k, self.L, self.task_prios)
            self.obs = real_build_obs(self.beta, self.prev, self.M, self.N)
        self.last_counts = cnt
        return rewards, self.done

    def pretransition(self):
        """real_constellation_env.py:232-244."""
        return {"beta": self.beta, "obs": self.obs, "prev_assigns": self.prev,
                "avail_actions": np.ones((self.B, self.n, self.m), dtype=bool)}


# ------------------------------------------------------------------ RealPowerConstellationEnv / InterferenceConstellationEnv
def power_update(power: np.ndarray, beta0_chosen: np.ndarray) -> np.ndarray:
    """real_power_constellation_env.py:172-180 (interference_constellation_env.py: same block): a live agent (power > 0)
    spends 0.2 on a task it can see (benefit > 1e-12) and recharges 0.1 (capped at 1) otherwise; a dead agent stays as it
    is.  float64 throughout: 1 - 5 * 0.2 leaves 5.55e-17, which is > 0 here but < 1e-12 in beta_hat."""
    alive = power > 0
    spent = power - 0.2
    charged = np.minimum(power + 0.1, 1.0)
    return np.where(alive, np.where(beta0_chosen > 1e-12, spent, charged), power)


class PowerState(RealState):
    """B RealPowerConstellationEnv instances (real_power_constellation_env.py:118-355): the real env plus a float64
    power state per agent, a zero reward / zero beta_hat for agents out of power, and N + 1 power values appended to every
    observation row.  ``prev0`` replaces the np.random.choice draw of reset (:130)."""

    def __init__(self, S, L, M, N, lambda_, task_prios=None, T_trans=None, T_ctor=None):
        super().__init__(S, L, M, N, lambda_, task_prios, T_trans, T_ctor)
        self.obs_size = real_obs_size(M, N, self.L) + N + 1                     # :291-293
        self.power = np.ones((self.B, self.n))

    def _obs(self):
        obs, _, nbrs, _ = real_build_obs(self.beta, self.prev, self.M, self.N, return_indices=True)
        bi = np.arange(self.B)[:, None, None]
        tail = np.concatenate([self.power[:, :, None], self.power[bi, nbrs]], axis=-1)   # :243-247
        return np.concatenate([obs, tail], axis=-1)

    def reset(self, prev0=None):
        self.k, self.done = 0, False
        self.beta = real_window(self.S, 0, self.L, self.task_prios)
        self.prev = (np.broadcast_to(np.arange(self.n), (self.B, self.n)) if prev0 is None else np.asarray(prev0)).copy()
        self.power = np.ones((self.B, self.n))                                    # :131
        self.obs = self._obs()
        return self.obs

    def _rewards(self, a, cnt):
        bh = real_beta_hat_chosen(self.beta, self.prev, a, self.lambda_, self.T_trans)
        bh = np.where(self.power < 1e-12, 0.0, bh)                                # :351-355 (beta_hat zeroed)
        return np.where(self.power > 0, _split_rewards(bh, a, cnt), 0.0)          # :157-165

    def step(self, actions):
        a = np.asarray(actions, dtype=np.int64)
        cnt = _counts(a, self.m)                                                  # every agent counts, dead or not (:147-149)
        rewards = self._rewards(a, cnt)
        bi = np.arange(self.B)[:, None]
        beta0 = self.beta[bi, np.arange(self.n)[None, :], a, 0]
        self.k += 1
        self.done = self.k >= self.T
        self.power = power_update(self.power, beta0)                              # :172-180
        self.prev = a.copy()
        if self.done:
            self.beta = np.zeros((self.B, self.n, self.m, self.L))
            self.obs = np.zeros((self.B, self.n, self.obs_size))
        else:
            self.beta = real_window(self.S, self.k, self.L, self.task_prios)
            self.obs = self._obs()
        self.last_counts = cnt
        return rewards, self.done

    def pretransition(self):
        out = super().pretransition()
        out["power_states"] = self.power.copy()                                   # :268
        return out


class InterferenceState(PowerState):
    """B InterferenceConstellationEnv instances: the power env whose reward is interference_reward_function
    (interference_constellation_env.py:309-353): agents in the same frequency band that serve neighbouring regions halve
    each other's benefit, only "applicable" agents (alive and on a task they can see) count for splitting, interference and
    the hand-over penalty."""

    def __init__(self, S, L, M, N, lambda_, neighbor_matrix, sat_freq_bands, task_prios=None):
        super().__init__(S, L, M, N, lambda_, task_prios)
        self.neighbor = np.asarray(neighbor_matrix, dtype=np.float64)
        self.bands = np.asarray(sat_freq_bands, dtype=np.int64)
        if self.bands.ndim == 1:
            self.bands = np.broadcast_to(self.bands, (self.B, self.n))

    def _rewards(self, a, cnt_all):
        B, n = self.B, self.n
        bi = np.arange(B)[:, None]
        beta0 = self.beta[bi, np.arange(n)[None, :], a, 0]
        applicable = ((self.power > 0) & ~(beta0 < 1e-12)).astype(np.float64)      # :314-316
        cnt = np.zeros((B, self.m))
        for b in range(B):
            np.add.at(cnt[b], a[b], applicable[b])                                # :318-321
        same_band = self.bands[:, :, None] == self.bands[:, None, :]              # [B, i, i']
        nb = self.neighbor[a[:, :, None], a[:, None, :]]                          # neighbor[a_i, a_i']
        conflicts = (nb * same_band * applicable[:, None, :]).sum(-1) - 1.0       # :331 (beams do not self-conflict)
        r = beta0 * 0.5 ** conflicts                                              # :332
        c = cnt[bi, a]
        r = np.where(c > 0, r / np.where(c > 0, c, 1.0), r)                      # :344-345
        r = np.where((applicable > 0) & (self.prev != a), r - self.lambda_, r)    # :348-349
        return r


# ------------------------------------------------------------------ MockConstellationEnv
def mock_obs(S: np.ndarray, k: int, L: int, curr_assignment: np.ndarray) -> np.ndarray:
    """obs_i = [curr_assignment[i,:] | S[i,:,k] | ... | S[i,:,k+L-1]] zero-padded past T
    (mock_constellation_env.py:107-112, 147-152)."""
    B, n, m, T = S.shape
    parts = [curr_assignment]
    for l in range(L):
        parts.append(S[..., k + l] if k + l < T else np.zeros((B, n, m)))
    return np.concatenate(parts, axis=-1)


class MockState:
    def __init__(self, S, L, lambda_, T_trans=None):
        self.S = np.asarray(S, dtype=np.float64)
        self.B, self.n, self.m, self.T = self.S.shape
        self.L, self.lambda_ = L, lambda_
        self.T_trans = None if T_trans is None else np.asarray(T_trans, dtype=np.float64)
        self.obs_size = (L + 1) * self.m
        self.k = 0

    def reset(self, prev_assigns):
        """mock_constellation_env.py:94-114.  ``prev_assigns`` replaces the reference's
        np.random.choice(m, n, replace=False) draw (:105) - injected for parity."""
        self.k = 0
        self.curr = np.zeros((self.B, self.n, self.m))
        self.beta = self.S[..., 0].copy()
        self.prev = np.asarray(prev_assigns, dtype=np.int64).copy()
        self.obs = mock_obs(self.S, 0, self.L, self.curr)
        return self.obs

    def step(self, actions):
        """mock_constellation_env.py:116-162."""
        a = np.asarray(actions, dtype=np.int64)
        B, n, m = self.B, self.n, self.m
        if self.T_trans is None:
            pen = (a != self.prev).astype(np.float64)
        else:
            pen = self.T_trans[self.prev, a]
        chosen = np.take_along_axis(self.beta, a[..., None], axis=2)[..., 0]
        bh = chosen - self.lambda_ * pen * (chosen > 1e-12)                # :263-270
        cnt = _counts(a, m)
        rewards = _split_rewards(bh, a, cnt)
        self.curr = np.zeros((B, n, m))
        np.put_along_axis(self.curr, a[..., None], 1.0, axis=2)
        self.k += 1
        self.obs = mock_obs(self.S, self.k, self.L, self.curr)            # post-done obs is NOT zero
        done = self.k >= self.T
        self.beta = self.S[..., self.k].copy() if not done else np.zeros((B, n, m))
        self.prev = a.copy()
        self.last_counts = cnt
        return rewards, done

    def pretransition(self):
        """mock_constellation_env.py:164-175 (no prev_assigns!)."""
        return {"obs": self.obs, "beta": self.beta,
                "avail_actions": np.ones((self.B, self.n, self.m), dtype=bool)}


# ------------------------------------------------------------------ action selectors
def random_available_action(avail: np.ndarray, u_action: np.ndarray) -> np.ndarray:
    """Contract for the injected-draw replacement of Categorical(avail.float()).sample()
    (classic_selectors.py:51): the floor(u * n_avail)-th available action, u in [0,1) fp32."""
    av = avail.astype(bool)
    n_av = av.sum(-1)
    u32 = np.asarray(u_action, dtype=np.float32)
    r = np.floor(u32 * n_av.astype(np.float32)).astype(np.int64)           # fp32 multiply, like the kernel
    r = np.minimum(r, np.maximum(n_av - 1, 0))
    csum = np.cumsum(av, axis=-1)                                           # rank of each available action
    hit = av & (csum == (r[..., None] + 1))
    return hit.argmax(-1).astype(np.int64)


def first_argmax(x: np.ndarray) -> np.ndarray:
    """tensor.max(dim)[1] -> first index on ties (verified on torch 2.11)."""
    return np.argmax(x, axis=-1).astype(np.int64)


def select_epsilon_greedy(q, avail, eps, u_explore, u_action):
    """EpsilonGreedyActionSelector.select_action (classic_selectors.py:37-54) with injected draws."""
    q = np.asarray(q, dtype=np.float32)
    masked = np.where(np.asarray(avail).astype(bool), q, np.float32(NEG_INF))   # :46-47
    pick_random = np.asarray(u_explore, dtype=np.float32) < np.float32(eps)     # :49-50 (fp32 compare)
    rnd = random_available_action(avail, u_action)                              # :51
    greedy = first_argmax(masked)
    return np.where(pick_random, rnd, greedy).astype(np.int64)                  # :53


def top_m_tasks(beta: np.ndarray, M: int) -> np.ndarray:
    """th.topk(beta.sum(-1), M) under the canonical stable rule (filtered_classic_selectors.py:50)."""
    tot = np.asarray(beta, dtype=np.float64).sum(-1)
    return np.argsort(-tot, axis=-1, kind="stable")[..., :M]


def select_filtered_epsilon_greedy(q, top, avail, m, eps, u_tie, u_explore, u_action):
    """FilteredEpsilonGreedyActionSelector.select_action (filtered_classic_selectors.py:17-63).

    q [B,n,M+1] fp32; top [B,n,M] = top-M task indices (descending); u_tie [B,n,m] fp32.
    row[j] = fl32(base + fl32(u_tie[j]*1e-8)); row[top[q]] = q[q]; greedy = first argmax over m.
    ``avail`` only affects the random branch (no -inf masking in this selector).
    """
    q = np.asarray(q, dtype=np.float32)
    base = q[..., -1]
    noise = (np.asarray(u_tie, dtype=np.float32) * np.float32(1e-8)).astype(np.float32)
    row = (base[..., None] + noise).astype(np.float32)                           # :45-47
    np.put_along_axis(row, np.asarray(top, dtype=np.int64), q[..., :-1], axis=-1)  # :53-54
    pick_random = np.asarray(u_explore, dtype=np.float32) < np.float32(eps)
    rnd = random_available_action(avail, u_action)
    greedy = first_argmax(row)
    return np.where(pick_random, rnd, greedy).astype(np.int64)


# ------------------------------------------------------------------ episode buffer semantics
def calc_conflicting_actions(actions: np.ndarray, m: int) -> float:
    """learners/q_learner.py:157-170, loop for loop."""
    B, T, n = actions.shape[:3]
    dup = 0
    for b in range(B):
        for k in range(T):
            cnt = np.zeros(m)
            for i in range(n):
                cnt[int(actions[b, k, i])] += 1
            dup += np.where(cnt > 0, cnt - 1, 0).sum()
    return dup / B / T


def calc_raw_benefits(beta: np.ndarray, actions: np.ndarray) -> float:
    """learners/q_learner.py:172-191 (4-D beta), loop for loop."""
    B, T, n = actions.shape[:3]
    tot = 0.0
    for b in range(B):
        for k in range(T):
            for i in range(n):
                tot += beta[b, k, i, int(actions[b, k, i])]
    return tot / B / T / n


def sample_categorical(probs: np.ndarray, u: np.ndarray, avail: np.ndarray | None = None) -> np.ndarray:
    """Contract of the policy-sampling selectors' Categorical(p).sample() (classic_selectors.py:25-26, :61-63) with an
    injected uniform per row: the first k with cdf[k] > u * cdf[-1], cdf in float64 (zero-probability actions are never
    drawn, like torch.multinomial)."""
    p = probs.astype(np.float64)
    if avail is not None:
        p = p * (np.asarray(avail) != 0)
    cdf = np.cumsum(p, axis=-1)
    target = u.astype(np.float64)[..., None] * cdf[..., -1:]
    return np.minimum((cdf <= target).sum(-1), p.shape[-1] - 1).astype(np.int64)


def sap_noise_std(benefit: np.ndarray, eps: float) -> np.ndarray:
    """stds = ones * mean|benefit[b]| * eps * 2, fp32 like torch (sap_selectors.py:84-85)."""
    avg = np.abs(benefit.astype(np.float32)).mean(axis=(1, 2), dtype=np.float32)
    return (avg * np.float32(eps) * np.float32(2)).astype(np.float32)


def lsa_maximize(benefit: np.ndarray, z: np.ndarray | None = None, std: np.ndarray | None = None):
    """Per env: scipy.optimize.linear_sum_assignment(benefit + z * std, maximize=True) (sap_selectors.py:77-91; the
    dependency is scipy, unpinned in the reference's requirements.txt:24).  Returns (col_ind [B, n], objective [B])."""
    from scipy.optimize import linear_sum_assignment

    b32 = benefit.astype(np.float32)
    if z is not None:
        b32 = (b32 + (z.astype(np.float32) * std.astype(np.float32)[:, None, None]).astype(np.float32)).astype(np.float32)
    cols, obj = [], []
    for b in range(b32.shape[0]):
        r, c = linear_sum_assignment(b32[b], maximize=True)
        cols.append(c)
        obj.append(b32[b][r, c].astype(np.float64).sum())
    return np.stack(cols).astype(np.int64), np.asarray(obj)


def haa_actions(beta: np.ndarray, prev: np.ndarray, lambda_: float, T_trans=None, buffer_dtype=np.float16) -> np.ndarray:
    """HAASelector.select_action (non_rl_selectors.py:18-50) on real-env state: beta is read back from the episode
    batch, i.e. rounded to the scheme dtype (fp16, real_constellation_env.py:96), then
    linear_sum_assignment(beta_hat(beta, prev)[..., 0], maximize=True) per env."""
    from scipy.optimize import linear_sum_assignment

    b = np.asarray(beta).astype(buffer_dtype).astype(np.float64)
    bh = real_beta_hat_full(b, np.asarray(prev, dtype=np.int64), lambda_, T_trans)[..., 0]
    return np.stack([linear_sum_assignment(bh[e], maximize=True)[1] for e in range(bh.shape[0])]).astype(np.int64)


def time_interval_sequences(L: int):
    """GENERATE_ALL_TIME_INTERVALS + BUILD_TIME_INTERVAL_SEQUENCES (utils/methods.py:309-349): every way to cut the next L
    steps into consecutive intervals, in the reference's enumeration order (depth first, shorter first interval first)."""
    out = []

    def rec(seq, start):
        if start == L:
            out.append(tuple(seq))
            return
        for end in range(start, L):
            rec(seq + [(start, end)], end + 1)

    rec([], 0)
    return out


def haal_actions(S, k, prev, L, lambda_, task_prios=None, T_trans=None):
    """HAALSelector.select_action (non_rl_selectors.py:54-118) for B real envs at step k: for every time-interval sequence,
    roll a copy of the env forward - per interval the optimal assignment of sum_l beta_hat(beta, prev)[..., l], held for
    the whole interval - and keep the first-interval assignment of the sequence with the largest summed reward (the first
    such sequence on ties).  Envs are float64 throughout, like the deep-copied reference envs."""
    from scipy.optimize import linear_sum_assignment

    S = np.asarray(S, dtype=np.float64)
    B, n, m, T = S.shape
    eff = min(L, T - k)
    picks = np.zeros((B, n), dtype=np.int64)
    for b in range(B):
        best_val, best = -np.inf, None
        for tis in time_interval_sequences(eff):
            kk, pv, val, first = k, np.asarray(prev[b], dtype=np.int64), 0.0, None
            for i, (t0, t1) in enumerate(tis):
                beta = real_window(S[b:b + 1], kk, L, task_prios)
                total = real_beta_hat_full(beta, pv[None], lambda_, T_trans).sum(-1)[0]
                a = linear_sum_assignment(total, maximize=True)[1]
                for _ in range(t1 - t0 + 1):
                    beta = real_window(S[b:b + 1], kk, L, task_prios)
                    bh = real_beta_hat_chosen(beta, pv[None], a[None], lambda_, T_trans)
                    val += float(_split_rewards(bh, a[None], _counts(a[None], m)).sum())
                    kk, pv = kk + 1, a
                if i == 0:
                    first = a
            if val > best_val:
                best_val, best = val, first
        picks[b] = best
    return picks


def filtered_benefit_matrix(q: np.ndarray, top: np.ndarray, m: int, u_tie: np.ndarray) -> np.ndarray:
    """filtered_sap_selectors.py:43-57: baseline + U * 1e-8 everywhere, the top-M tasks get their own Q-values (fp32)."""
    B, n, _ = q.shape
    mat = (np.broadcast_to(q[:, :, -1:], (B, n, m)).astype(np.float32)
           + (u_tie.astype(np.float32) * np.float32(1e-8)).astype(np.float32)).astype(np.float32)
    np.put_along_axis(mat, top.astype(np.int64), q[:, :, :-1].astype(np.float32), axis=2)
    return mat


def one_hot(actions: np.ndarray, m: int, dtype) -> np.ndarray:
    """OneHot.transform then cast to the actions dtype (transforms.py:16-19, episode_buffer.py:123-129)."""
    out = np.zeros(actions.shape + (m,), dtype=dtype)
    np.put_along_axis(out, actions[..., None].astype(np.int64), 1, axis=-1)
    return out


def rollout(state, policy, scheme_kind: str, prev0=None):
    """EpisodeRunner.run loop order (episode_runner.py:60-127) over a batch of envs.

    ``policy(t, pre) -> actions [B,n]``.  Returns a dict of float64/int64 arrays shaped
    [B, T+1, ...] holding what the episode buffer receives BEFORE dtype casting
    (A.5 timeline): obs, beta, prev_assigns (real only), actions, rewards, terminated, filled.
    """
    B, n, m, T = state.B, state.n, state.m, state.T
    if scheme_kind == "real" and not isinstance(state, PowerState):
        state.reset()            # the real env starts from prev_assigns = arange(n) (:129)
    else:
        state.reset(prev0)       # mock and power-type envs draw prev_assigns (injected)
    pre = state.pretransition()
    out = {
        "obs": np.zeros((B, T + 1, n, state.obs_size)),
        "beta": np.zeros((B, T + 1) + pre["beta"].shape[1:]),
        "actions": np.zeros((B, T + 1, n), dtype=np.int64),
        "rewards": np.zeros((B, T + 1, n)),
        "terminated": np.zeros((B, T + 1), dtype=bool),
        "filled": np.zeros((B, T + 1), dtype=np.int64),
        "counts": np.zeros((B, T + 1, m), dtype=np.int64),
    }
    if scheme_kind == "real":
        out["prev_assigns"] = np.zeros((B, T + 1, n), dtype=np.int64)
    if "power_states" in pre:
        out["power_states"] = np.zeros((B, T + 1, n))
    t, done = 0, False
    while not done:
        pre = state.pretransition()
        out["obs"][:, t], out["beta"][:, t], out["filled"][:, t] = pre["obs"], pre["beta"], 1
        if scheme_kind == "real":
            out["prev_assigns"][:, t] = pre["prev_assigns"]
        if "power_states" in pre:
            out["power_states"][:, t] = pre["power_states"]
        a = np.asarray(policy(t, pre), dtype=np.int64)
        r, done = state.step(a)
        out["actions"][:, t], out["rewards"][:, t], out["terminated"][:, t] = a, r, done
        out["counts"][:, t] = state.last_counts
        t += 1
    pre = state.pretransition()
    out["obs"][:, t], out["beta"][:, t], out["filled"][:, t] = pre["obs"], pre["beta"], 1
    if scheme_kind == "real":
        out["prev_assigns"][:, t] = pre["prev_assigns"]
    if "power_states" in pre:
        out["power_states"][:, t] = pre["power_states"]
    return out


# ------------------------------------------------------------------ synthetic benefits
def gen_dense(rng: np.random.Generator, B, n, m, T) -> np.ndarray:
    """G-dense: U(0,1) fp32-representable values, returned as [B,n,m,T] float32."""
    return rng.random((B, n, m, T), dtype=np.float32)


def gen_exact(rng: np.random.Generator, B, n, m, T, zero_frac: float = 0.0) -> np.ndarray:
    """G-exact: k*2^-10 grid (k in [0,1024)) so that every L-sum is exact in fp32 and fp64.
    ``zero_frac`` > 0 makes a tie-heavy variant (exact zeros and frequent duplicates)."""
    k = rng.integers(0, 1024, size=(B, n, m, T)).astype(np.float32)
    S = k * np.float32(2.0 ** -10)
    if zero_frac > 0:
        S = np.where(rng.random((B, n, m, 1)) < zero_frac, np.float32(0), S)
    return S.astype(np.float32)


def gen_ref_like(rng: np.random.Generator, B, n, m, T, p_active=0.25, wmin=3.0, wmax=6.0) -> np.ndarray:
    """G-ref: the law of generate_benefits_over_time (mock_constellation_env.py:276-299), vectorised,
    rounded to fp32 (P(active)=0.25, Gaussian bump in time, per-task scale in {1,1,1,10})."""
    scale = rng.choice(np.array([1.0, 1.0, 1.0, 10.0]), size=(B, 1, m, 1))
    active = rng.random((B, n, m, 1)) > (1.0 - p_active)
    center = rng.uniform(0, T, size=(B, n, m, 1))
    spread = rng.uniform(wmin, wmax, size=(B, n, m, 1))
    sigma_2 = np.sqrt(spread ** 2 / -8 / np.log(0.05))
    t = np.arange(T, dtype=np.float64)[None, None, None, :]
    S = scale * np.exp(-(t - center) ** 2 / sigma_2 / 2) * active
    return S.astype(np.float32)
