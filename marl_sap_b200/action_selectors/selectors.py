"""Action selectors backed by the sm_100a selection kernels.

EpsilonGreedyActionSelector          <- /root/reference/src/action_selectors/classic_selectors.py:28-54
FilteredEpsilonGreedyActionSelector  <- /root/reference/src/action_selectors/filtered_classic_selectors.py:6-63

Same constructor / ``select_action(agent_inputs, avail_actions, t_env, test_mode=False, beta=None)``
signature and ``.epsilon`` / ``.schedule`` attributes.  Random draws come from an in-kernel Philox4x32-10
stream keyed by (seed, env*n+agent, episode counter, step) or, for parity tests, from injected uniforms
(``inject_draws``).  Inputs must be CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import torch as th

from .. import _lib
from ..components.epsilon_schedules import DecayThenFlatSchedule


class _KernelSelectorBase:
    def __init__(self, args):
        self.args = args
        self.schedule = DecayThenFlatSchedule(args.epsilon_start, args.epsilon_finish, args.epsilon_anneal_time,
                                              decay="linear")
        self.epsilon = self.schedule.eval(0)
        self.seed = self._base_seed = int(getattr(args, "seed", 0) or 0) & 0xFFFFFFFFFFFFFFFF
        self.env_offset = 0
        self.envs = None            # assigned by the runners (reference: episode_runner.py:39)
        self.episode_ctr = None     # device uint64 scalar (as int64 tensor), set by the runner
        self.step_k = None          # device int32 [B] step counters of the batched env
        self._injected = None
        self._calls = 0
        self._local_ctr = None

    def set_env_offset(self, env_offset):
        """Global index of this process's first env (multi-GPU runs: rank r owns envs [offset, offset + B)).  The kernels
        key Philox by (seed; local row, episode, step); folding the offset into the seed keeps the streams of different
        ranks apart (identical seeds on every rank would make all ranks explore identically).  Offset 0 leaves the seed
        as given."""
        self.env_offset = int(env_offset)
        if self.env_offset:
            z = (self._base_seed + 0x9E3779B97F4A7C15 * (self.env_offset + 1)) & 0xFFFFFFFFFFFFFFFF  # splitmix64 finaliser
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
            self.seed = z ^ (z >> 31)
        else:
            self.seed = self._base_seed

    def bind_counters(self, episode_ctr, step_k):
        """Give the kernel device-side (episode, step) counters so each (env, agent, t) draw is unique and the
        launch stays CUDA-graph capturable."""
        self.episode_ctr, self.step_k = episode_ctr, step_k

    def use_device_epsilon(self, enabled=True):
        """Keep epsilon in a device scalar the kernels read (needed when select_action is captured in a CUDA graph:
        the schedule keeps moving between replays).  ``set_device_epsilon`` refreshes it outside the graph."""
        self._eps_dev_enabled = bool(enabled)

    def _eps_tensor(self, device):
        # allocated once: captured graphs keep its address
        if getattr(self, "_eps_dev", None) is None:
            self._eps_dev = th.zeros(1, dtype=th.float32, device=device)
        return self._eps_dev

    def set_device_epsilon(self, t_env, test_mode, device):
        eps = self._eps(t_env, test_mode)
        self._eps_tensor(device).fill_(eps)
        return eps

    def _eps_ptr(self, eps, device):
        if not getattr(self, "_eps_dev_enabled", False):
            return None
        if getattr(self, "_eps_dev", None) is None:
            self._eps_tensor(device).fill_(eps)
        return self._eps_dev.data_ptr()

    def inject_draws(self, **draws):
        """Parity hook: u_explore/u_action[/u_tie] fp32 tensors consumed by the next select_action call."""
        self._injected = draws

    def _eps(self, t_env, test_mode):
        self.epsilon = self.schedule.eval(t_env)
        if test_mode:
            self.epsilon = self.args.evaluation_epsilon
        return float(self.epsilon)

    def _counters(self, device, B):
        """Counters for callers that did not bind any: a private device counter advanced per call."""
        if self.episode_ctr is not None:
            return self.episode_ctr, self.step_k
        if self._local_ctr is None or self._local_ctr.device != device:
            self._local_ctr = th.zeros(1, dtype=th.int64, device=device)
        else:
            self._local_ctr += 1
        return self._local_ctr, None

    @staticmethod
    def _avail_u8(avail_actions):
        if avail_actions is None:
            return None
        av = avail_actions
        if av.dim() == 3 and av.stride(0) == 0 and av.stride(1) == 0 and av.stride(2) == 0:
            return None  # constant all-ones view of a lazy EpisodeBatch: every action available, nothing to read
        if av.dtype == th.bool:
            av = av.contiguous().view(th.uint8)
        elif av.dtype != th.uint8:
            av = (av != 0).contiguous().view(th.uint8)
        return av.contiguous()


class EpsilonGreedyActionSelector(_KernelSelectorBase):
    graph_capturable = True  # one kernel, all draws and counters on the device

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, beta=None):
        eps = self._eps(t_env, test_mode)
        _lib.require_cuda(agent_inputs, "agent_inputs")
        q = agent_inputs.detach()
        if q.dtype != th.float32:
            q = q.float()
        q = q.contiguous()
        B, n, A = q.shape
        av = self._avail_u8(avail_actions)
        if av is not None and av.stride(0) == 0:
            av = None  # constant all-ones view from a lazy EpisodeBatch: everything available
        out = th.empty(B, n, dtype=th.int64, device=q.device)
        inj = self._injected or {}
        self._injected = None
        ue, ua = inj.get("u_explore"), inj.get("u_action")
        ctr, k = self._counters(q.device, B)
        if k is not None and k.numel() != B:
            k = None
        lib = _lib.load()
        _lib.check(lib.sap_select_epsilon_greedy(q.data_ptr(), _lib.ptr(av), B, n, A, eps, self._eps_ptr(eps, q.device),
                                                 self.seed, _lib.ptr(ctr),
                                                 _lib.ptr(k), _lib.ptr(ue), _lib.ptr(ua), out.data_ptr(),
                                                 _lib.stream_ptr(q.device)), "sap_select_epsilon_greedy")
        return out


    def fused_select_args(self, agent_inputs, t_env, test_mode=False):
        """The same selection as ``select_action`` (every action available), described for ``sap_rollout_step``: the env
        kernel then selects and steps in one launch.  Returns (SapSelectArgs, actions_out [B, n] int64, keepalive)."""
        eps = self._eps(t_env, test_mode)
        _lib.require_cuda(agent_inputs, "agent_inputs")
        q = agent_inputs.detach()
        if q.dtype != th.float32:
            q = q.float()
        q = q.contiguous()
        B, n, _ = q.shape
        out = th.empty(B, n, dtype=th.int64, device=q.device)
        inj = self._injected or {}
        self._injected = None
        ue, ua = inj.get("u_explore"), inj.get("u_action")
        ctr, _ = self._counters(q.device, B)
        a = _lib.SapSelectArgs()
        a.q, a.eps_dev, a.episode_ctr = q.data_ptr(), self._eps_ptr(eps, q.device), _lib.ptr(ctr)
        a.u_explore, a.u_action, a.seed, a.eps = _lib.ptr(ue), _lib.ptr(ua), self.seed, eps
        return a, out, (q, ue, ua, ctr)


class FilteredEpsilonGreedyActionSelector(_KernelSelectorBase):
    """Epsilon-greedy in the "top-M tasks + anything-else baseline" action space.

    ``top`` (int32 [B, n, M], the env's own top-M task indices) can be passed instead of ``beta``; when only
    ``beta`` is given the top-M is recomputed from it with the stable rule (``sap_topm_from_beta``).
    """

    graph_capturable = True

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, beta=None, top=None):
        assert beta is not None or top is not None, "Need beta to figure out which are the top M tasks for each agent."
        eps = self._eps(t_env, test_mode)
        _lib.require_cuda(agent_inputs, "agent_inputs")
        q = agent_inputs.detach()
        if q.dtype != th.float32:
            q = q.float()
        q = q.contiguous()
        B, n, Mp1 = q.shape
        M = self.args.env_args["M"]
        assert Mp1 == M + 1, f"agent must output M+1={M + 1} values, got {Mp1}"
        lib = _lib.load()
        stream = _lib.stream_ptr(q.device)
        if top is None:
            _lib.require_cuda(beta, "beta")
            bt = beta.contiguous()
            if bt.dtype not in (th.float32, th.float16):
                bt = bt.float()
            m, L = bt.shape[2], (bt.shape[3] if bt.dim() == 4 else 1)
            top = th.empty(B, n, M, dtype=th.int32, device=q.device)
            _lib.check(lib.sap_topm_from_beta(bt.data_ptr(), _lib.sap_dtype(bt.dtype), B, n, m, L, M, top.data_ptr(), stream),
                       "sap_topm_from_beta")
        else:
            top = top.contiguous()
            m = avail_actions.shape[2] if avail_actions is not None else self.args.env_args["m"]
        av = self._avail_u8(avail_actions)
        if av is not None and av.stride(0) == 0:
            av = None
        out = th.empty(B, n, dtype=th.int64, device=q.device)
        inj = self._injected or {}
        self._injected = None
        ut, ue, ua = inj.get("u_tie"), inj.get("u_explore"), inj.get("u_action")
        ctr, k = self._counters(q.device, B)
        if k is not None and k.numel() != B:
            k = None
        _lib.check(lib.sap_select_filtered_epsilon_greedy(q.data_ptr(), top.data_ptr(), _lib.ptr(av), B, n, m, M, eps,
                                                          self._eps_ptr(eps, q.device), self.seed, _lib.ptr(ctr), _lib.ptr(k), _lib.ptr(ut), _lib.ptr(ue),
                                                          _lib.ptr(ua), out.data_ptr(), stream),
                   "sap_select_filtered_epsilon_greedy")
        return out
