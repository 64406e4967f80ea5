"""Device-resident batches of sequential-assignment environments (B independent envs per launch).

``BatchedRealConstellationEnv``  <- /root/reference/src/envs/real_constellation_env.py  (RealConstellationEnv :17-328)
``BatchedMockConstellationEnv``  <- /root/reference/src/envs/mock_constellation_env.py  (MockConstellationEnv :13-274)

State lives in HBM: benefit planes [B, T, n, m] fp32 (re-laid out once from the reference's
[n, m, T]), step counters k[B], prev_assigns[B, n], per-env return accumulators.  ``reset`` and
``step`` are single kernel launches that write their slots straight into an EpisodeBatch.
There is no CPU path: construction fails without CUDA or without the built library.
"""
from __future__ import annotations

import numpy as np
import torch as th

from .. import _lib
from ..components.transforms import OneHot


def real_obs_size(M, N, L):
    """real_constellation_env.py:263-265."""
    return M * L + N * M * L + N * M // 2 * L + M


def real_scheme(n, m, L, obs_size):
    """real_constellation_env.py:89-114 (discrete actions)."""
    scheme = {
        "obs": {"vshape": obs_size, "group": "agents", "dtype": th.float16},
        "actions": {"vshape": (1,), "group": "agents", "dtype": th.int16},
        "avail_actions": {"vshape": (m,), "group": "agents", "dtype": th.bool},
        "rewards": {"vshape": (n,), "dtype": th.float16},
        "terminated": {"vshape": (1,), "dtype": th.bool},
        "prev_assigns": {"vshape": (n,), "dtype": th.int16, "part_of_state": True},
        "beta": {"vshape": (n, m, L), "dtype": th.float16, "part_of_state": True},
    }
    return scheme, {"actions": ("actions_onehot", [OneHot(out_dim=m)])}


def mock_scheme(n, m, L):
    """mock_constellation_env.py:67-92 (discrete actions)."""
    scheme = {
        "obs": {"vshape": L * m + m, "group": "agents", "dtype": th.float32},
        "actions": {"vshape": (1,), "group": "agents", "dtype": th.int64},
        "avail_actions": {"vshape": (m,), "group": "agents", "dtype": th.bool},
        "rewards": {"vshape": (n,), "dtype": th.float32},
        "terminated": {"vshape": (1,), "dtype": th.bool},
        "prev_assigns": {"vshape": (n,), "dtype": th.int64, "part_of_state": True},
        "beta": {"vshape": (n, m), "dtype": th.float32, "part_of_state": True},
    }
    return scheme, {"actions": ("actions_onehot", [OneHot(out_dim=m)])}


class _BatchedEnvBase:
    kind = "base"

    def __init__(self, B, n, m, T, L, M, N, lambda_, T_trans, device):
        self.lib = _lib.load()
        if not th.cuda.is_available():
            raise RuntimeError("marl_sap_b200: a CUDA device is required (there is no CPU path)")
        self.device = th.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError(f"marl_sap_b200: envs live on a CUDA device, got {self.device}")
        self.B, self.n, self.m, self.T, self.L, self.M, self.N = B, n, m, T, L, M, N
        self.lambda_ = float(lambda_)
        dev = self.device
        self.k = th.zeros(B, dtype=th.int32, device=dev)
        self.prev = th.zeros(B, n, dtype=th.int32, device=dev)
        self.ep_return = th.zeros(B, dtype=th.float64, device=dev)
        self.counts = th.zeros(B, m, dtype=th.int32, device=dev)
        self.T_trans = None if T_trans is None else th.as_tensor(np.asarray(T_trans), dtype=th.float32).contiguous().to(dev)
        if self.T_trans is not None and tuple(self.T_trans.shape) != (m, m):
            raise ValueError(f"T_trans must be [m, m] = [{m}, {m}], got {tuple(self.T_trans.shape)}")
        self.planes = None
        self.plane_stats = None
        self.shared_planes = False
        self._staging = None
        self._benefit_source = None  # planes by generation, for replay buffers that rebuild `beta` lazily
        self.t_host = 0  # host mirror of k (all envs step in lockstep)
        self.launches_per_step = 1  # kernels one reset()/step() call enqueues (4 on the multi-CTA path of large shapes)

    def __deepcopy__(self, memo):
        """A batched env is device state (benefit planes, counters, the loaded library): callers that deep-copy an
        object graph holding one - the reference's learners do ``copy.deepcopy(mac)`` for their target network, and the
        MAC's jump-start selector is bound to the runner's env - get the SAME env back, not a second copy of the planes."""
        return self

    def benefit_source(self):
        """The ``BenefitSource`` through which episode batches with a lazy ``beta`` find the planes they were rolled out on
        (components/episode_buffer.py).  Planes a source holds are never overwritten in place: ``load_benefits`` /
        ``generate_benefits`` then write a fresh tensor, so that stored episodes keep reading the benefits they saw."""
        if self._benefit_source is None:
            from ..components.episode_buffer import BenefitSource

            self._benefit_source = BenefitSource(self.kind, self.n, self.m, self.T, self.L)
        return self._benefit_source

    def register_planes(self):
        """(source, generation) of the planes the next episode runs on."""
        src = self.benefit_source()
        return src, src.register(self.planes, getattr(self, "task_prios", None), self.shared_planes)

    def _planes_held(self):
        return self._benefit_source is not None and self.planes is not None and self._benefit_source.holds(self.planes)

    # ------------------------------------------------------------------ benefits
    def load_benefits(self, sat_prox_mat):
        """Install benefit tensors given in the REFERENCE layout: [n, m, T] (shared by all B envs, like
        ParallelRunner's identical env_args) or [B, n, m, T].  Host arrays are uploaded with
        ``sap_benefit_upload_host``; CUDA tensors are re-laid out in place with ``sap_benefit_ingest``."""
        S = sat_prox_mat
        if isinstance(S, np.ndarray):
            S64 = S if S.dtype == np.float64 else None
            S = th.from_numpy(np.ascontiguousarray(S, dtype=np.float32))
            if S64 is not None and not np.array_equal(S.numpy().astype(np.float64), S64):
                self._warn_fp32_benefits(float(np.max(np.abs(S.numpy().astype(np.float64) - S64))))
        if S.dtype != th.float32:
            S32 = S.to(th.float32)
            if S.dtype == th.float64 and not th.equal(S32.double(), S):
                self._warn_fp32_benefits(float((S32.double() - S).abs().max()))
            S = S32
        shared = S.dim() == 3
        Bp = 1 if shared else S.shape[0]
        if tuple(S.shape[-3:]) != (self.n, self.m, self.T) or (not shared and Bp != self.B):
            raise ValueError(f"sat_prox_mat must be [n,m,T]=[{self.n},{self.m},{self.T}] or [B,n,m,T] with B={self.B}, "
                             f"got {tuple(S.shape)}")
        S = S.contiguous()
        if self.planes is None or self.planes.shape[0] != Bp or self._planes_held():
            self.planes = th.empty(Bp, self.T, self.n, self.m, dtype=th.float32, device=self.device)
        stream = _lib.stream_ptr(self.device)
        if S.is_cuda:
            _lib.check(self.lib.sap_benefit_ingest(S.data_ptr(), self.planes.data_ptr(), Bp, self.n, self.m, self.T, stream),
                       "sap_benefit_ingest")
            self._keepalive = S
        else:
            if self._staging is None or self._staging.numel() != S.numel():
                self._staging = th.empty(S.numel(), dtype=th.float32, device=self.device)
            _lib.check(self.lib.sap_benefit_upload_host(S.data_ptr(), self._staging.data_ptr(), self.planes.data_ptr(), Bp,
                                                        self.n, self.m, self.T, stream), "sap_benefit_upload_host")
            self._keepalive = S  # the async copy reads the host buffer until the stream reaches it
        self.shared_planes = shared
        self._refresh_plane_stats()
        return self

    def _warn_fp32_benefits(self, max_err):
        """The planes are fp32 (half the HBM traffic of the step kernel).  The reference computes on float64: benefits that
        fp32 cannot represent are ROUNDED here, which can reorder near-tied top-M / rival lists and the `> 1e-12` tests
        against a float64 run.  Every fixture of the test-suite is fp32-representable; real simulator output is not."""
        import warnings

        warnings.warn(f"marl_sap_b200: sat_prox_mat holds float64 values that fp32 cannot represent (max rounding error "
                      f"{max_err:.3g}); the device envs compute on the fp32-rounded benefits, so rankings of values closer "
                      "than that may differ from a float64 reference run", RuntimeWarning, stacklevel=3)

    def _refresh_plane_stats(self):
        """Per-plane {min, max} metadata (sap_benefit_stats): lets the real-env kernel scale its selection keys
        without a second pass over the window.  One pass over the planes per episode."""
        Bp = self.planes.shape[0]
        if getattr(self, "plane_stats", None) is None or self.plane_stats.shape[0] != Bp:
            self.plane_stats = th.empty(Bp, self.T, 2, dtype=th.float32, device=self.device)
        _lib.check(self.lib.sap_benefit_stats(self.planes.data_ptr(), self.plane_stats.data_ptr(), Bp, self.n, self.m,
                                              self.T, _lib.stream_ptr(self.device)), "sap_benefit_stats")

    def set_planes(self, planes, shared=False, stats=None):
        """Adopt an already device-laid-out tensor [B or 1, T, n, m] (no copy)."""
        _lib.require_cuda(planes, "planes")
        assert planes.dtype == th.float32 and planes.is_contiguous()
        assert tuple(planes.shape[1:]) == (self.T, self.n, self.m)
        self.planes, self.shared_planes = planes, bool(shared)
        if stats is not None:
            self.plane_stats = stats
        else:
            self._refresh_plane_stats()
        return self

    def dims(self):
        return _lib.SapEnvDims(self.B, self.n, self.m, self.T, self.L, self.M, self.N, 1 if self.shared_planes else 0)

    def _check_batch(self, batch):
        if batch.batch_size != self.B or batch.max_seq_length != self.T + 1:
            raise ValueError(f"EpisodeBatch must be [B={self.B}, T+1={self.T + 1}], got "
                             f"[{batch.batch_size}, {batch.max_seq_length}]")
        if self.planes is None:
            raise RuntimeError("no benefit tensor loaded: call load_benefits() first")

    def _check_actions(self, actions):
        _lib.require_cuda(actions, "actions")
        if actions.dtype != th.int64:
            actions = actions.long()
        actions = actions.reshape(self.B, self.n).contiguous()
        return actions

    def enable_bids_as_actions(self):
        """``bids_as_actions`` of the reference envs (real_constellation_env.py:75-80, 110-112, 140-141;
        mock_constellation_env.py:53-56, 88-90, 121-122): an action is a bid per task, the buffer's ``actions`` field holds
        the bids (fp32 [m]), and ``step`` turns every env's bid matrix into an assignment with one batched
        linear-sum-assignment before the usual step."""
        self.bids_as_actions = True
        self.scheme["actions"] = {"vshape": (self.m,), "group": "agents", "dtype": th.float32}
        self.preprocess = {}

    def _actions_for_step(self, actions, view, batch):
        if not getattr(self, "bids_as_actions", False):
            return self._check_actions(actions)
        from ..action_selectors.sap_selectors import lsa_maximize

        _lib.require_cuda(actions, "actions")
        bids = actions.detach().float().reshape(self.B, self.n, self.m).contiguous()
        batch.data.transition_data["actions"][:, self.t_host] = bids   # the kernel must not write its int actions there
        view.actions = _lib.SapField()
        return lsa_maximize(bids)

    def episode_returns(self):
        return self.ep_return


class BatchedRealConstellationEnv(_BatchedEnvBase):
    kind = "real"

    def __init__(self, B, n, m, T, L, M, N, lambda_, sat_prox_mat=None, task_prios=None, T_trans=None, device=None,
                 T_ctor=None):
        # L = min(L, T_ctor) is taken BEFORE T is overridden by sat_prox_mat.shape[2]
        # (real_constellation_env.py:38 vs :58-60); callers pass T_ctor when they mirror the ctor.
        L = min(L, T_ctor if T_ctor is not None else T)
        super().__init__(B, n, m, T, L, M, N, lambda_, T_trans, device)
        self.obs_size = real_obs_size(M, N, L)
        self.task_prios = None if task_prios is None else \
            th.as_tensor(np.asarray(task_prios), dtype=th.float32).contiguous().to(self.device)
        self.top = th.zeros(B, n, M, dtype=th.int32, device=self.device)
        self._top_next, self._ahead = None, False
        self._top_home = self.top   # reset() always starts from this buffer (captured CUDA graphs bake the alternation in)
        self.scheme, self.preprocess = real_scheme(n, m, L, self.obs_size)
        need = int(self.lib.sap_real_scratch_doubles(self.dims()))
        self.scratch = th.empty(need, dtype=th.float64, device=self.device) if need > 0 else None
        self.launches_per_step = 4 if need > 0 else 1  # prep + keys + lists + main (csrc/sap_real_large.cu)
        if sat_prox_mat is not None:
            self.load_benefits(sat_prox_mat)

    def reset(self, batch):
        self._check_batch(batch)
        if self.top is not self._top_home:
            self.top, self._top_next = self._top_next, self.top
        self._ahead = False
        view = batch.kernel_view()
        _lib.check(self.lib.sap_real_reset(self.dims(), _lib.ptr(self.planes), _lib.ptr(self.plane_stats),
                                           _lib.ptr(self.task_prios), _lib.ptr(self.k), _lib.ptr(self.prev),
                                           _lib.ptr(self.ep_return), view, _lib.ptr(self.top), _lib.ptr(self.scratch),
                                           _lib.stream_ptr(self.device)), "sap_real_reset")
        self.t_host = 0

    # ------------------------------------------------------------------ observation ahead of the step
    def supports_obs_ahead(self, batch):
        """True when the observation of slot t + 1 can be built next to the agent forward of step t
        (``sap_real_obs_ahead``): the shipped configuration on the one-CTA-per-env kernel."""
        return (type(self) is BatchedRealConstellationEnv and self.task_prios is None and self.plane_stats is not None
                and not getattr(self, "bids_as_actions", False) and batch.scheme["obs"]["dtype"] == th.float16
                and bool(self.lib.sap_real_obs_ahead_ok(self.dims())))

    @staticmethod
    def _with_agent_in(view, agent_in):
        if agent_in is not None:
            f = _lib.SapField()
            f.ptr, f.env_stride, f.t_stride, f.dtype = agent_in.data_ptr(), agent_in.stride(0), agent_in.stride(1), _lib.sap_dtype(agent_in.dtype)
            view.agent_in = f
        return view

    def obs_ahead(self, batch, agent_in=None):
        """Enqueue (on the current stream) the observation build of slot k + 1: it reads the benefit window only, so it may
        run while the agent network and the selector are still working on slot k.  The following ``step`` then only
        computes rewards / counters and sets the "previous task in my top-M" flags.  ``agent_in``: the staging rows to fill
        (the runner double-buffers them: the agent is reading the other buffer)."""
        view = self._with_agent_in(batch.kernel_view(), agent_in)
        if self._top_next is None:
            self._top_next = th.zeros_like(self.top)
        _lib.check(self.lib.sap_real_obs_ahead(self.dims(), _lib.ptr(self.planes), _lib.ptr(self.plane_stats), _lib.ptr(self.k),
                                               view, _lib.ptr(self._top_next), _lib.stream_ptr(self.device)), "sap_real_obs_ahead")
        self._ahead = True

    def step(self, actions, batch, agent_in=None):
        view = self._with_agent_in(batch.kernel_view(), agent_in)
        actions = self._actions_for_step(actions, view, batch)
        if self._ahead:  # the observation rows of the new slot are already there
            _lib.check(self.lib.sap_real_step_after_obs(self.dims(), _lib.ptr(self.planes), _lib.ptr(self.task_prios),
                                                        _lib.ptr(self.T_trans), self.lambda_, actions.data_ptr(), _lib.ptr(self.k),
                                                        _lib.ptr(self.prev), _lib.ptr(self.ep_return), _lib.ptr(self.counts), view,
                                                        _lib.ptr(self._top_next), _lib.stream_ptr(self.device)),
                       "sap_real_step_after_obs")
            self._ahead = False
            self.t_host += 1
            if self.t_host < self.T:  # env.top = top-M tasks of the newest observation (a finished env builds none)
                self.top, self._top_next = self._top_next, self.top
            return self.t_host >= self.T
        _lib.check(self.lib.sap_real_step(self.dims(), _lib.ptr(self.planes), _lib.ptr(self.plane_stats),
                                          _lib.ptr(self.task_prios),
                                          _lib.ptr(self.T_trans), self.lambda_, actions.data_ptr(), _lib.ptr(self.k),
                                          _lib.ptr(self.prev), _lib.ptr(self.ep_return), _lib.ptr(self.counts), view,
                                          _lib.ptr(self.top), _lib.ptr(self.scratch), _lib.stream_ptr(self.device)),
                   "sap_real_step")
        self.t_host += 1
        return self.t_host >= self.T

    def supports_select_step(self, batch):
        """True when ``step_select`` applies (``sap_rollout_step``: selection + step in one launch)."""
        return self.supports_obs_ahead(batch) and bool(self.lib.sap_rollout_step_ok(self.dims()))

    def step_select(self, sel, actions_out, batch, agent_in=None):
        """Epsilon-greedy selection (``sel``: a ``SapSelectArgs`` from the selector's ``fused_select_args``) and the env
        step in one launch; the chosen actions land in ``actions_out`` [B, n] int64.  Same results as
        ``selector.select_action`` + ``step``."""
        view = self._with_agent_in(batch.kernel_view(), agent_in)
        ahead = self._ahead
        _lib.check(self.lib.sap_rollout_step(sel, self.dims(), _lib.ptr(self.planes), _lib.ptr(self.plane_stats),
                                             _lib.ptr(self.T_trans), self.lambda_, actions_out.data_ptr(), _lib.ptr(self.k),
                                             _lib.ptr(self.prev), _lib.ptr(self.ep_return), _lib.ptr(self.counts), view,
                                             _lib.ptr(self.top), _lib.ptr(self._top_next) if ahead else None,
                                             _lib.stream_ptr(self.device)), "sap_rollout_step")
        self._ahead = False
        self.t_host += 1
        if ahead and self.t_host < self.T:
            self.top, self._top_next = self._top_next, self.top
        return self.t_host >= self.T

    def beta_field(self, dtype=th.float16):
        """The `beta` buffer field [B, T+1, n, m, L] rebuilt from the planes (lazy materialisation)."""
        out = th.empty(self.B, self.T + 1, self.n, self.m, self.L, dtype=dtype, device=self.device)
        _lib.check(self.lib.sap_real_beta_window(self.dims(), _lib.ptr(self.planes), _lib.ptr(self.task_prios),
                                                 out.data_ptr(), _lib.sap_dtype(dtype), _lib.stream_ptr(self.device)),
                   "sap_real_beta_window")
        return out


def power_scheme(n, m, L, obs_size):
    """real_power_constellation_env.py:95-116 / interference_constellation_env.py:70-98: the real scheme + power_states."""
    scheme, preprocess = real_scheme(n, m, L, obs_size)
    scheme["power_states"] = {"vshape": (n,), "dtype": th.float16, "part_of_state": True}
    return scheme, preprocess


class BatchedRealPowerConstellationEnv(BatchedRealConstellationEnv):
    """B RealPowerConstellationEnv instances (/root/reference/src/envs/real_power_constellation_env.py:17-358).

    One step = ``sap_power_pre`` (who is out of power, then the float64 power update) -> ``sap_real_step_ex`` (the real
    env's step + observation kernel with a zero reward for dead agents, the rival indices exported and N + 1 free columns
    behind every observation row) -> ``sap_power_post`` (power columns, ``power_states`` field).  ``reset`` takes the
    ``prev_assigns`` draw (``np.random.choice(m, n, replace=False)``, :130) as ``prev0`` like the mock env."""

    kind = "real"

    def __init__(self, B, n, m, T, L, M, N, lambda_, sat_prox_mat=None, task_prios=None, T_trans=None, device=None, T_ctor=None):
        super().__init__(B, n, m, T, L, M, N, lambda_, sat_prox_mat=sat_prox_mat, task_prios=task_prios, T_trans=T_trans,
                         device=device, T_ctor=T_ctor)
        if self.scratch is not None:
            raise NotImplementedError(f"power / interference envs run on the one-CTA-per-env kernel: n={n}, m={m} does not fit "
                                      "one SM's shared memory (the multi-CTA path has no power columns yet)")
        self.base_obs_size = self.obs_size
        self.obs_size = self.base_obs_size + N + 1                      # :291-293
        self.scheme, self.preprocess = power_scheme(n, m, self.L, self.obs_size)
        self.power = th.ones(B, n, dtype=th.float64, device=self.device)
        self.dead = th.zeros(B, n, dtype=th.uint8, device=self.device)
        self.nbr = th.zeros(B, n, N, dtype=th.int32, device=self.device)
        self.launches_per_step = 3

    def draw_prev_assigns(self):
        if self.m < self.n:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        return np.stack([np.random.choice(self.m, self.n, replace=False) for _ in range(self.B)]).astype(np.int64)

    def _post(self, batch, view):
        pf = _lib.field_of(batch.data.transition_data["power_states"]) if "power_states" in batch.data.transition_data \
            else _lib.SapField()
        _lib.check(self.lib.sap_power_post(self.dims(), _lib.ptr(self.k), _lib.ptr(self.power), _lib.ptr(self.nbr), view, pf,
                                           self.base_obs_size, self.obs_size, _lib.stream_ptr(self.device)), "sap_power_post")

    def reset(self, batch, prev0=None):
        self._check_batch(batch)
        if prev0 is None:
            prev0 = self.draw_prev_assigns()
        prev0 = th.as_tensor(np.asarray(prev0), dtype=th.int64).reshape(self.B, self.n).contiguous().to(self.device)
        self.power.fill_(1.0)                                            # :131
        view = batch.kernel_view()
        _lib.check(self.lib.sap_real_reset_ex(self.dims(), _lib.ptr(self.planes), _lib.ptr(self.plane_stats),
                                              _lib.ptr(self.task_prios), _lib.ptr(self.k), _lib.ptr(self.prev),
                                              _lib.ptr(self.ep_return), view, _lib.ptr(self.top), prev0.data_ptr(),
                                              _lib.ptr(self.nbr), self.obs_size, _lib.stream_ptr(self.device)),
                   "sap_real_reset_ex")
        self._post(batch, view)
        self.t_host = 0

    def _rewards_before_step(self, actions, view, batch):
        """Power env: the env kernel computes the rewards itself, with the dead mask."""
        return self.dead, self.ep_return

    def step(self, actions, batch):
        view = batch.kernel_view()
        actions = self._actions_for_step(actions, view, batch)
        stream = _lib.stream_ptr(self.device)
        dead, ep_return = self._rewards_before_step(actions, view, batch)
        _lib.check(self.lib.sap_power_pre(self.dims(), _lib.ptr(self.planes), _lib.ptr(self.task_prios), actions.data_ptr(),
                                          _lib.ptr(self.k), _lib.ptr(self.power), _lib.ptr(self.dead), stream), "sap_power_pre")
        _lib.check(self.lib.sap_real_step_ex(self.dims(), _lib.ptr(self.planes), _lib.ptr(self.plane_stats),
                                             _lib.ptr(self.task_prios), _lib.ptr(self.T_trans), self.lambda_,
                                             actions.data_ptr(), _lib.ptr(self.k), _lib.ptr(self.prev), _lib.ptr(ep_return),
                                             _lib.ptr(self.counts), view, _lib.ptr(self.top), _lib.ptr(dead),
                                             _lib.ptr(self.nbr), self.obs_size, stream), "sap_real_step_ex")
        self._post(batch, view)
        self.t_host += 1
        return self.t_host >= self.T


class BatchedInterferenceConstellationEnv(BatchedRealPowerConstellationEnv):
    """B InterferenceConstellationEnv instances (/root/reference/src/envs/interference_constellation_env.py:17-406): the
    power env whose reward is ``interference_reward_function`` (:309-353, ``sap_interference_rewards``).  The reference builds
    its proximity tensor and region neighbour matrix from the orbit simulator (out of scope); here they are arguments."""

    def __init__(self, B, n, m, T, L, M, N, lambda_, sat_prox_mat, neighbor_matrix, sat_freq_bands, task_prios=None,
                 device=None, T_ctor=None):
        super().__init__(B, n, m, T, L, M, N, lambda_, sat_prox_mat=sat_prox_mat, task_prios=task_prios, device=device,
                         T_ctor=T_ctor)
        nb = th.as_tensor(np.asarray(neighbor_matrix), dtype=th.float32)
        if tuple(nb.shape) != (m, m):
            raise ValueError(f"neighbor_matrix must be [m, m] = [{m}, {m}], got {tuple(nb.shape)}")
        self.neighbor = nb.contiguous().to(self.device)
        bands = th.as_tensor(np.asarray(sat_freq_bands), dtype=th.int32)
        if tuple(bands.shape) not in ((n,), (B, n)):
            raise ValueError(f"sat_freq_bands must be [n] or [B, n], got {tuple(bands.shape)}")
        self.bands_per_env = bands.dim() == 2
        self.bands = bands.contiguous().to(self.device)
        self._scratch_return = th.zeros(B, dtype=th.float64, device=self.device)
        self.launches_per_step = 4

    def _rewards_before_step(self, actions, view, batch):
        rf = view.rewards
        _lib.check(self.lib.sap_interference_rewards(self.dims(), _lib.ptr(self.planes), _lib.ptr(self.task_prios),
                                                     _lib.ptr(self.neighbor), _lib.ptr(self.bands), int(self.bands_per_env),
                                                     self.lambda_, actions.data_ptr(), _lib.ptr(self.k), _lib.ptr(self.prev),
                                                     _lib.ptr(self.power), _lib.ptr(self.ep_return), rf,
                                                     _lib.stream_ptr(self.device)), "sap_interference_rewards")
        view.rewards = _lib.SapField()   # the env kernel must not overwrite them; its own return goes to a scratch buffer
        return None, self._scratch_return

    def reset(self, batch, prev0=None):
        super().reset(batch, prev0)
        self.ep_return.zero_()


class BatchedMockConstellationEnv(_BatchedEnvBase):
    kind = "mock"

    def __init__(self, B, n, m, T, L, lambda_, sat_prox_mat=None, T_trans=None, device=None, generate_seed=None):
        super().__init__(B, n, m, T, L, 0, 0, lambda_, T_trans, device)
        self.obs_size = (L + 1) * m
        self.scheme, self.preprocess = mock_scheme(n, m, L)
        # constant_benefits = False in the reference (no sat_prox_mat given, mock_constellation_env.py:32-37): benefits are
        # drawn in the constructor (widths 5..8) and redrawn at every reset (widths 3..6, :99-100) - here on the device
        self.constant_benefits = generate_seed is None
        self._gen_seed, self._gen_episode = generate_seed, 0
        if sat_prox_mat is not None:
            self.load_benefits(sat_prox_mat)
        elif generate_seed is not None:
            self.generate_benefits(5.0, 8.0)

    def generate_benefits(self, width_min, width_max):
        """``generate_benefits_over_time`` for all B envs at once, straight into the planes (``sap_benefit_generate``)."""
        if self.planes is None or self.planes.shape[0] != self.B or self._planes_held():
            self.planes = th.empty(self.B, self.T, self.n, self.m, dtype=th.float32, device=self.device)
            self.shared_planes = False
        _lib.check(self.lib.sap_benefit_generate(self.planes.data_ptr(), self.B, self.n, self.m, self.T, float(width_min),
                                                 float(width_max), int(self._gen_seed or 0) & 0xFFFFFFFFFFFFFFFF,
                                                 self._gen_episode, _lib.stream_ptr(self.device)), "sap_benefit_generate")
        self._gen_episode += 1

    def draw_prev_assigns(self):
        """mock_constellation_env.py:105: np.random.choice(m, n, replace=False) per env, host RNG like the reference."""
        if self.m < self.n:
            raise ValueError("Cannot take a larger sample than population when 'replace=False'")
        return np.stack([np.random.choice(self.m, self.n, replace=False) for _ in range(self.B)]).astype(np.int64)

    def reset(self, batch, prev0=None):
        self._check_batch(batch)
        if not self.constant_benefits:
            self.generate_benefits(3.0, 6.0)  # :99-100
        if prev0 is None:
            prev0 = self.draw_prev_assigns()
        prev0 = th.as_tensor(np.asarray(prev0) if not isinstance(prev0, th.Tensor) else prev0, dtype=th.int64)
        prev0 = prev0.reshape(self.B, self.n).contiguous().to(self.device)
        view = batch.kernel_view()
        _lib.check(self.lib.sap_mock_reset(self.dims(), _lib.ptr(self.planes), prev0.data_ptr(), _lib.ptr(self.k),
                                           _lib.ptr(self.prev), _lib.ptr(self.ep_return), view,
                                           _lib.stream_ptr(self.device)), "sap_mock_reset")
        self.t_host = 0

    def step(self, actions, batch):
        view = batch.kernel_view()
        actions = self._actions_for_step(actions, view, batch)
        _lib.check(self.lib.sap_mock_step(self.dims(), _lib.ptr(self.planes), _lib.ptr(self.T_trans), self.lambda_,
                                          actions.data_ptr(), _lib.ptr(self.k), _lib.ptr(self.prev),
                                          _lib.ptr(self.ep_return), _lib.ptr(self.counts), view,
                                          _lib.stream_ptr(self.device)), "sap_mock_step")
        self.t_host += 1
        return self.t_host >= self.T
