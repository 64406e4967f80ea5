"""Single-environment objects with the reference's gym-style env API, backed by the batched CUDA envs.

These are what ``envs.REGISTRY[name](**env_args)`` returns, so ``run.py``-style callers and
EpisodeRunner-style loops keep working:  attrs n, m, T, L, scheme, preprocess;  reset();
step(actions) -> (list rewards, done, {});  get_pretransition_data() -> dict of 1-element lists.
(/root/reference/src/envs/real_constellation_env.py:17-328, mock_constellation_env.py:13-274.)

Each call launches the same kernels as the batched runners with B = 1 and reads the slot back to
the host, so this path is API-compatible but slow; throughput comes from runners.ParallelRunner.
Objects are picklable / deep-copyable: device state is rebuilt on first use.
"""
from __future__ import annotations

import copy

import numpy as np
import torch as th

from ..components.episode_buffer import EpisodeBatch
from .batched import (BatchedInterferenceConstellationEnv, BatchedMockConstellationEnv, BatchedRealConstellationEnv,
                      BatchedRealPowerConstellationEnv, mock_scheme, power_scheme, real_obs_size, real_scheme)


def generate_benefits_over_time(n, m, T, width_min, width_max, scale_min=0.25, scale_max=2):
    """Host generator with the same law and the same numpy RNG call order as
    mock_constellation_env.py:276-299 (device generation is a "next" row)."""
    benefits = np.zeros((n, m, T))
    t = np.arange(T, dtype=np.float64)
    for j in range(m):
        benefit_scale = np.random.choice([1, 1, 1, 10])
        for i in range(n):
            task_active = 1 if np.random.rand() > 0.75 else 0
            if task_active:
                time_center = np.random.uniform(0, T)
                time_spread = np.random.uniform(width_min, width_max)
                sigma_2 = np.sqrt(time_spread ** 2 / -8 / np.log(0.05))
                benefits[i, j, :] = benefit_scale * np.exp(-(t - time_center) ** 2 / sigma_2 / 2)
    return benefits


class _SingleEnvBase:
    _kind = None

    # ------------------------------------------------------------------ lazily built device state
    def _device_state(self):
        if self._impl is None:
            self._impl = self._make_impl()
            if self.bids_as_actions:
                self._impl.enable_bids_as_actions()
            scheme = copy.deepcopy(self.scheme)
            for key in ("obs", "rewards", "beta"):  # read back at fp32, returned to the caller as float64
                scheme[key]["dtype"] = th.float32
            groups = {"agents": self.n}
            self._batch = EpisodeBatch(scheme, groups, 1, self.T + 1, preprocess=self.preprocess, device=self._impl.device)
        return self._impl, self._batch

    def __getstate__(self):
        st = self.__dict__.copy()
        st["_impl"] = None
        st["_batch"] = None
        return st

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        st = self.__getstate__()
        new.__dict__.update({k: copy.deepcopy(v, memo) for k, v in st.items()})
        if self._impl is not None:  # carry the live state over (HAAL rolls copies forward)
            impl, batch = new._device_state()
            impl.k.copy_(self._impl.k)
            impl.prev.copy_(self._impl.prev)
            impl.t_host = self._impl.t_host
            for k, v in self._batch.data.transition_data.items():
                batch.data.transition_data[k].copy_(v)
        return new

    # ------------------------------------------------------------------ reference env API
    def step(self, actions):
        impl, batch = self._device_state()
        t = self.k
        if self.bids_as_actions:  # a bid per task and agent; the assignment is solved on the device (:140-141)
            a = th.as_tensor(np.asarray(actions, dtype=np.float32), device=impl.device).reshape(1, self.n, self.m)
        else:
            a = th.as_tensor(np.asarray(actions, dtype=np.int64), device=impl.device).reshape(1, self.n)
        done = impl.step(a, batch)
        self.k += 1
        self.done = bool(done)
        rewards = batch["rewards"][0, t].double().cpu().numpy()
        self.prev_assigns = impl.prev[0].cpu().numpy().astype(int) if self.bids_as_actions else np.asarray(actions, dtype=int)
        self._refresh(batch)
        return [r for r in rewards], self.done, {}

    def _refresh(self, batch):
        t = self.k
        self._obs = [o for o in batch["obs"][0, t].double().cpu().numpy()]
        self.beta = batch["beta"][0, t].double().cpu().numpy()

    def close(self):
        return True

    def seed(self, seed=None):
        self._seed = seed

    def get_obs(self):
        return self._obs

    def get_obs_agent(self, agent_id):
        return self._obs[agent_id]

    def get_obs_size(self):
        return self.obs_space_size

    def get_avail_actions(self):
        return [self.get_avail_agent_actions(i) for i in range(self.n)]

    def get_avail_agent_actions(self, agent_id):
        return [1] * self.get_total_actions()

    def get_total_actions(self):
        return self.m

    def get_stats(self):
        return {}

    def render(self):
        raise NotImplementedError

    def save_replay(self):
        raise NotImplementedError


class RealConstellationEnv(_SingleEnvBase):
    _kind = "real"

    def __init__(self, num_planes, num_sats_per_plane, m, T, N, M, L, lambda_, sat_prox_mat=None, graphs=None,
                 bids_as_actions=False, seed=None, T_trans=None, task_prios=None, device=None):
        self.seed(seed)
        if sat_prox_mat is None or graphs is None:
            raise NotImplementedError("orbit-propagated benefits (HighPerformanceConstellationSim, poliastro) are out of "
                                      "scope: pass sat_prox_mat=[n,m,T] and graphs (real_constellation_env.py:47-60)")
        self.N, self.M = N, M
        self.L = min(L, T)  # :38 - clamped with the ctor's T, before T is overridden below
        self.constant_benefits = True
        self.sat_prox_mat = np.asarray(sat_prox_mat)
        self.graphs = graphs
        self.n, self.m, self.T = self.sat_prox_mat.shape  # :58-60
        self.lambda_ = lambda_
        self.T_trans = T_trans
        self.task_prios_arg = task_prios
        self.bids_as_actions = bool(bids_as_actions)
        self.k, self.done, self.beta, self.prev_assigns, self._obs = 0, False, None, None, None
        self.obs_space_size = real_obs_size(self.M, self.N, self.L)
        self.scheme, self.preprocess = real_scheme(self.n, self.m, self.L, self.obs_space_size)
        if self.bids_as_actions:  # :110-112
            self.scheme["actions"] = {"vshape": (self.m,), "group": "agents", "dtype": th.float32}
            self.preprocess = {}
        self._device = device
        self._impl = self._batch = None

    def _make_impl(self):
        return BatchedRealConstellationEnv(1, self.n, self.m, self.T, self.L, self.M, self.N, self.lambda_,
                                           sat_prox_mat=self.sat_prox_mat, task_prios=self.task_prios_arg,
                                           T_trans=self.T_trans, device=self._device)

    def reset(self):
        impl, batch = self._device_state()
        impl.reset(batch)
        self.k, self.done = 0, False
        self.prev_assigns = np.arange(self.n)
        self._refresh(batch)
        return self.get_obs()

    def get_pretransition_data(self):
        """real_constellation_env.py:232-244."""
        return {"beta": [self.beta], "obs": [self._obs], "prev_assigns": [self.prev_assigns],
                "avail_actions": [self.get_avail_actions()]}

    def beta_hat(self, beta, prev_assigns):
        """real_constellation_env.py:282-328 as array ops (not on the rollout hot path: the step kernel
        evaluates beta_hat only at the chosen entries).  Accepts an optional leading time dimension."""
        beta = beta.cpu().numpy() if isinstance(beta, th.Tensor) else np.asarray(beta)
        prev = prev_assigns.cpu().numpy() if isinstance(prev_assigns, th.Tensor) else np.asarray(prev_assigns)
        squeeze = beta.ndim == 3
        if squeeze:
            beta = beta[None]
        if prev.ndim == 1:
            prev = prev[None]
        Tt = np.ones((self.m, self.m)) - np.eye(self.m) if self.T_trans is None else np.asarray(self.T_trans)
        pen = Tt[prev.astype(np.int64)]
        out = beta.astype(np.float64, copy=True)
        out[..., 0] = out[..., 0] - self.lambda_ * (pen * (beta.sum(-1) > 1e-12))
        return out[0] if squeeze else out


class RealPowerConstellationEnv(RealConstellationEnv):
    """/root/reference/src/envs/real_power_constellation_env.py:17-358 behind the same facade: per-agent power states
    (float64 on the device), N + 1 power columns per observation, ``power_states`` in the pre-transition data."""

    def __init__(self, num_planes, num_sats_per_plane, m, T, N, M, L, lambda_, sat_prox_mat=None, graphs=None,
                 bids_as_actions=False, seed=None, T_trans=None, task_prios=None, device=None):
        if bids_as_actions:
            raise NotImplementedError("bids_as_actions is built for the real and mock envs only")
        super().__init__(num_planes, num_sats_per_plane, m, T, N, M, L, lambda_, sat_prox_mat=sat_prox_mat, graphs=graphs,
                         seed=seed, T_trans=T_trans, task_prios=task_prios, device=device)
        if task_prios is None:   # :70: a quarter of the tasks are high priority, drawn with the global numpy RNG
            self.task_prios_arg = np.random.choice([1, 1, 1, 5], size=self.m, replace=True).astype(np.float64)
        self.power_states = np.ones(self.n)
        self.obs_space_size = real_obs_size(self.M, self.N, self.L) + self.N + 1
        self.scheme, self.preprocess = power_scheme(self.n, self.m, self.L, self.obs_space_size)

    def _make_impl(self):
        return BatchedRealPowerConstellationEnv(1, self.n, self.m, self.T, self.L, self.M, self.N, self.lambda_,
                                                sat_prox_mat=self.sat_prox_mat, task_prios=self.task_prios_arg,
                                                T_trans=self.T_trans, device=self._device)

    def reset(self):
        impl, batch = self._device_state()
        self.prev_assigns = np.random.choice(self.m, self.n, replace=False)   # :130
        impl.reset(batch, prev0=self.prev_assigns[None])
        self.k, self.done = 0, False
        self._refresh(batch)
        return self.get_obs()

    def _refresh(self, batch):
        super()._refresh(batch)
        self.power_states = self._impl.power[0].cpu().numpy().copy()

    def get_pretransition_data(self):
        out = super().get_pretransition_data()
        out["power_states"] = [self.power_states]                              # :268
        return out

    def beta_hat(self, beta, prev_assigns, power_states):
        """:310-355: the real env's beta_hat, zeroed for agents whose power is below 1e-12."""
        out = super().beta_hat(beta, prev_assigns)
        pw = power_states.cpu().numpy() if isinstance(power_states, th.Tensor) else np.asarray(power_states)
        squeeze = out.ndim == 3
        if squeeze:
            out = out[None]
        pw = pw.reshape(out.shape[0], self.n)
        out = np.where((pw < 1e-12)[:, :, None, None], 0.0, out)
        return out[0] if squeeze else out


class InterferenceConstellationEnv(RealPowerConstellationEnv):
    """/root/reference/src/envs/interference_constellation_env.py:17-406.  The reference derives its proximity tensor and
    the region neighbour matrix from the orbit simulator (poliastro / h3: out of scope); this facade takes them as
    ``sat_prox_mat`` and ``neighbor_matrix`` and keeps the rest of the constructor (``res`` is accepted and ignored)."""

    def __init__(self, num_planes, num_sats_per_plane, res, T, N, M, L, lambda_, task_prios=None, sat_freq_bands=None,
                 bids_as_actions=False, seed=None, sat_prox_mat=None, neighbor_matrix=None, device=None):
        if sat_prox_mat is None or neighbor_matrix is None:
            raise NotImplementedError("coverage-task proximities and the region neighbour matrix come from the orbit simulator "
                                      "(HighPerformanceConstellationSim, out of scope): pass sat_prox_mat=[n,m,T] and "
                                      "neighbor_matrix=[m,m] (interference_constellation_env.py:54-61)")
        m = np.asarray(sat_prox_mat).shape[1]
        self.beam_types = 7
        self.constant_setup = task_prios is not None and sat_freq_bands is not None                 # :69-76
        if not self.constant_setup:
            task_prios = np.random.choice([1, 1, 1, 5], size=m, replace=True)
            sat_freq_bands = np.random.choice(list(range(self.beam_types)), size=np.asarray(sat_prox_mat).shape[0], replace=True)
        self.res = res
        self.neighbor_matrix = np.asarray(neighbor_matrix, dtype=np.float64)
        self.sat_freq_bands = np.asarray(sat_freq_bands)
        super().__init__(num_planes, num_sats_per_plane, m, T, N, M, L, lambda_, sat_prox_mat=sat_prox_mat, graphs=[1],
                         bids_as_actions=bids_as_actions, seed=seed, task_prios=np.asarray(task_prios, dtype=np.float64),
                         device=device)

    def _make_impl(self):
        return BatchedInterferenceConstellationEnv(1, self.n, self.m, self.T, self.L, self.M, self.N, self.lambda_,
                                                   self.sat_prox_mat, self.neighbor_matrix, self.sat_freq_bands,
                                                   task_prios=self.task_prios_arg, device=self._device)

    def reset(self):
        if not self.constant_setup:                                            # :132-141: new priorities and bands per episode
            self.task_prios_arg = np.random.choice([1, 1, 1, 5], size=self.m, replace=True).astype(np.float64)
            self.sat_freq_bands = np.random.choice(list(range(self.beam_types)), size=self.n, replace=True)
            self._impl = None
        return super().reset()


class MockConstellationEnv(_SingleEnvBase):
    _kind = "mock"

    def __init__(self, n, m, T, L, lambda_, bids_as_actions=False, seed=None, sat_prox_mat=None, T_trans=None,
                 device=None):
        self.seed(seed)
        self.n, self.m, self.T, self.L, self.lambda_ = n, m, T, L, lambda_
        if sat_prox_mat is None:
            self.constant_benefits = False
            self.sat_prox_mat = generate_benefits_over_time(n, m, T, 5, 8)  # :34
        else:
            self.constant_benefits = True
            self.sat_prox_mat = np.asarray(sat_prox_mat)
        self.T_trans = T_trans
        self.graphs = None
        self.bids_as_actions = bool(bids_as_actions)
        self.k, self.beta, self.prev_assigns, self._obs = 0, None, None, None
        self.obs_space_size = self.L * self.m + self.m
        self.scheme, self.preprocess = mock_scheme(n, m, L)
        if self.bids_as_actions:  # :88-90
            self.scheme["actions"] = {"vshape": (self.m,), "group": "agents", "dtype": th.float32}
            self.preprocess = {}
        self._device = device
        self._impl = self._batch = None

    def _make_impl(self):
        return BatchedMockConstellationEnv(1, self.n, self.m, self.T, self.L, self.lambda_, sat_prox_mat=self.sat_prox_mat,
                                           T_trans=self.T_trans, device=self._device)

    def reset(self):
        impl, batch = self._device_state()
        if not self.constant_benefits:
            self.sat_prox_mat = generate_benefits_over_time(self.n, self.m, self.T, 3, 6)  # :100
            impl.load_benefits(self.sat_prox_mat)
        self.prev_assigns = np.random.choice(self.m, self.n, replace=False)  # :105
        impl.reset(batch, prev0=self.prev_assigns[None])
        self.k = 0
        self._refresh(batch)
        return self.get_obs(), self.get_state()

    def get_state(self):
        return np.concatenate(self._obs, axis=0).astype(np.float32)

    def get_state_size(self):
        return self.n * self.get_obs_size()

    def get_env_info(self):
        return {"state_shape": self.get_state_size(), "obs_shape": self.get_obs_size(), "m": self.get_total_actions(),
                "n": self.n, "T": self.T}

    def get_pretransition_data(self):
        """mock_constellation_env.py:164-175 (no prev_assigns, like the reference)."""
        return {"obs": [self._obs], "avail_actions": [self.get_avail_actions()], "beta": [self.beta]}

    def beta_hat(self, beta, prev_assigns):
        """mock_constellation_env.py:228-274 as array ops (see RealConstellationEnv.beta_hat)."""
        beta = beta.cpu().numpy() if isinstance(beta, th.Tensor) else np.asarray(beta)
        prev = prev_assigns.cpu().numpy() if isinstance(prev_assigns, th.Tensor) else np.asarray(prev_assigns)
        squeeze = beta.ndim == 2
        if squeeze:
            beta = beta[None]
        if prev.ndim == 1:
            prev = prev[None]
        Tt = np.ones((self.m, self.m)) - np.eye(self.m) if self.T_trans is None else np.asarray(self.T_trans)
        pen = Tt[prev.astype(np.int64)]
        out = beta.astype(np.float64) - self.lambda_ * (pen * (beta > 1e-12))
        return out[0] if squeeze else out
