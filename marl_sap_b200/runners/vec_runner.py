"""Batched CUDA rollout runner behind the reference's runner interface.

Replaces /root/reference/src/runners/parallel_runner.py:12-299 (ParallelRunner + env_worker processes +
pickle over Pipes) and runners/episode_runner.py:8-137 (EpisodeRunner) for the SAP envs: all
``batch_size_run`` environments live in one device-resident batched env and advance in lockstep.  Per timestep: the torch
agent forward, one selection kernel and one env kernel (step -> obs(t+1) -> buffer writes; four launches on the multi-CTA
path of the large shapes).  For the shipped real-env configuration at the bench shape the default schedule is instead: the
observation of slot t + 1 built on a second stream next to the agent forward (``overlap_obs_build``) and ONE launch that
selects and steps (``fuse_select_step``, ``sap_rollout_step``); the T-step loop can be replayed as a CUDA graph
(``use_cuda_graph``).

Same surface as the reference runners: ``Runner(args, logger)``, ``setup(scheme, groups, preprocess, mac)``,
``get_env()``, ``run(test_mode) -> EpisodeBatch``, ``close_env()``, ``save_replay()``, attributes
``batch_size, t_env, T, log_train_stats_t`` and the ``logger.log_stat`` keys ``return_mean/std``,
``ep_length_mean``, ``epsilon``, ``steps`` (with the ``test_`` prefix in test mode).

Semantics follow EpisodeRunner (A.5 timeline of SURVEY.md): ``terminated`` is the env's done flag (True only
at t = T-1), no extra selection at t = T, independent per-(env, agent, step) random streams.  The reference
ParallelRunner's bugs (terminated list-truthiness, extra select, shared numpy RNG) are NOT
reproduced.

Multi-GPU: one process per GPU.  By default every rank owns ``batch_size_run`` envs (weak scaling); with
``args.batch_size_run_is_global = True`` the ``batch_size_run`` envs are block-partitioned over the ranks
(``utils.dist.env_partition``; strong scaling, the north star's "4096 envs sharded across 2/4/8 GPUs") and a
``sat_prox_mat`` of shape [batch_size_run, n, m, T] is sliced to the rank's block.  Either way a rank's random streams
are keyed by the GLOBAL index of its first env, so no two ranks draw the same exploration noise.  The only collective
is one all-reduce of [sum return, sum return^2, n_episodes, sum ep_length] per ``run()``.

``compat_parallel_runner_quirks = True`` reproduces what the reference ParallelRunner writes into ``terminated``
(parallel_runner.py:181-187: the flag comes from list truthiness instead of the env's done flag - env 0 is never marked
terminated, every other env always) and its extra selection stored at t = T; off by default, EpisodeRunner semantics
otherwise (golden ``runner_parallel.npz``).
"""
from __future__ import annotations

import math
from functools import partial

import numpy as np
import torch as th

from ..components.episode_buffer import EpisodeBatch
from ..envs import BATCHED as batched_REGISTRY
from ..envs import REGISTRY as env_REGISTRY
from ..utils.dist import env_partition


def build_batched_env(env_name, env_args, B, device):
    """Translate the reference's ``env_args`` (config yaml + explicit_dict_items) into a batched device env."""
    if env_name not in batched_REGISTRY:
        env_REGISTRY[env_name]()  # raises the explanatory NotImplementedError for off-path envs
    ea = dict(env_args)
    S = ea.get("sat_prox_mat", None)
    if S is None and env_name == "mock_constellation_env":
        # the reference's MockConstellationEnv draws its own benefits (ctor and every reset): generated on the device
        env = batched_REGISTRY[env_name](B, ea["n"], ea["m"], ea["T"], ea["L"], ea["lambda_"], T_trans=ea.get("T_trans"),
                                         device=device, generate_seed=int(ea.get("seed", 0) or 0) + 1)
        if ea.get("bids_as_actions"):
            env.enable_bids_as_actions()
        return env
    if S is None:
        raise NotImplementedError("batched real-env runners need env_args['sat_prox_mat'] ([n,m,T] shared or [B,n,m,T]): "
                                  "the orbit simulator that would produce it is out of scope (DESIGN.md)")
    shape = tuple(S.shape)
    n, m, T = shape[-3:]
    if env_name in ("real_constellation_env", "real_power_constellation_env"):
        prios = ea.get("task_prios")
        if prios is None and env_name == "real_power_constellation_env":   # real_power_constellation_env.py:70
            prios = np.random.choice([1, 1, 1, 5], size=m, replace=True).astype(np.float64)
        env = batched_REGISTRY[env_name](B, n, m, T, ea["L"], ea["M"], ea["N"], ea["lambda_"], sat_prox_mat=S,
                                         task_prios=prios, T_trans=ea.get("T_trans"), device=device,
                                         T_ctor=ea.get("T", T))
    elif env_name == "interference_constellation_env":
        if ea.get("neighbor_matrix") is None or ea.get("task_prios") is None or ea.get("sat_freq_bands") is None:
            raise NotImplementedError("the batched interference env needs env_args['neighbor_matrix'], ['task_prios'] and "
                                      "['sat_freq_bands'] (the reference derives / redraws them per env on the host)")
        env = batched_REGISTRY[env_name](B, n, m, T, ea["L"], ea["M"], ea["N"], ea["lambda_"], S, ea["neighbor_matrix"],
                                         ea["sat_freq_bands"], task_prios=ea["task_prios"], device=device, T_ctor=ea.get("T", T))
    else:
        env = batched_REGISTRY[env_name](B, n, m, T, ea["L"], ea["lambda_"], sat_prox_mat=S, T_trans=ea.get("T_trans"),
                                         device=device)
    if ea.get("bids_as_actions"):
        env.enable_bids_as_actions()
    return env


class _EnvInfo:
    """What ``run.py`` reads from ``runner.get_env()`` (run.py:113-131): shapes, scheme and preprocess."""

    def __init__(self, env):
        self.n, self.m, self.T, self.L = env.n, env.m, env.T, env.L
        self.scheme, self.preprocess = env.scheme, env.preprocess
        self.obs_size = env.obs_size
        self.M, self.N = getattr(env, "M", None), getattr(env, "N", None)

    def get_obs_size(self):
        return self.obs_size

    def get_total_actions(self):
        return self.m

    def get_env_info(self):
        return {"obs_shape": self.obs_size, "m": self.m, "n": self.n, "T": self.T}


class CudaVecRunner:
    """One batched env of ``batch_size_run`` environments per process (= per GPU)."""

    def __init__(self, args, logger):
        self.args = args
        self.logger = logger
        dev = getattr(args, "device", "cuda")
        self.device = th.device(dev if str(dev).startswith("cuda") else "cuda")
        rank, world = 0, 1
        if th.distributed.is_available() and th.distributed.is_initialized():
            rank, world = th.distributed.get_rank(), th.distributed.get_world_size()
        self.rank, self.world = rank, world
        self.global_batch = bool(getattr(args, "batch_size_run_is_global", False)) and world > 1
        if self.global_batch:  # strong scaling: this rank's block of the batch_size_run envs
            self.env_offset, self.batch_size = env_partition(self.args.batch_size_run, rank, world)
            if self.batch_size == 0:
                raise ValueError(f"batch_size_run={self.args.batch_size_run} leaves rank {rank} of {world} without an env")
        else:
            self.env_offset, self.batch_size = rank * self.args.batch_size_run, self.args.batch_size_run
        env_args = dict(self.args.env_args)
        S = env_args.get("sat_prox_mat", None)
        if self.global_batch and S is not None and len(S.shape) == 4:
            if S.shape[0] != self.args.batch_size_run:
                raise ValueError(f"sat_prox_mat has {S.shape[0]} envs, batch_size_run is {self.args.batch_size_run}")
            env_args["sat_prox_mat"] = S[self.env_offset:self.env_offset + self.batch_size]
        if env_args.get("sat_prox_mat", None) is None and "seed" in env_args:
            env_args["seed"] = int(env_args["seed"] or 0) + 7919 * self.env_offset  # device-generated benefits differ per rank
        self.env = build_batched_env(self.args.env, env_args, self.batch_size, self.device)
        self.compat_quirks = bool(getattr(args, "compat_parallel_runner_quirks", False))
        self.T = self.env.T
        self.t = 0
        self.t_env = 0
        self.lazy = tuple(getattr(args, "lazy_buffer_fields", ()) or ())
        self.reuse_batch = bool(getattr(args, "reuse_episode_batch", False))
        self.train_returns, self.test_returns = _Moments(), _Moments()
        self.train_stats, self.test_stats = {}, {}
        self.log_train_stats_t = -100000
        self.episode_ctr = th.zeros(1, dtype=th.int64, device=self.device)
        self.last_episode_returns = None
        self._batch = None
        self.kernel_launches = 0
        self.max_graphs = int(getattr(args, "max_cuda_graphs", 16))

    # ------------------------------------------------------------------ reference runner API
    def setup(self, scheme, groups, preprocess, mac):
        self.new_batch = partial(EpisodeBatch, scheme, groups, self.batch_size, self.T + 1, preprocess=preprocess,
                                 device=self.device, lazy=self.lazy)
        self.mac = mac
        self.scheme, self.groups, self.preprocess = scheme, groups, preprocess
        self.mac.action_selector.envs = [self.get_env()]
        if hasattr(self.mac.action_selector, "bind_counters"):
            self.mac.action_selector.bind_counters(self.episode_ctr, self.env.k)
        if hasattr(self.mac.action_selector, "set_env_offset"):
            self.mac.action_selector.set_env_offset(self.env_offset)
        js = getattr(self.mac, "jumpstart_action_selector", None)
        if js is not None:  # JumpstartMAC: the non-learning policy reads the env state directly
            js.envs = [self.get_env()]
            if hasattr(js, "bind_env"):
                js.bind_env(self.env)
        # fp32 staging of the agent network's input: the env kernel writes float(obs) of the new slot into it, so
        # the per-step `batch["obs"][:, t].float()` conversion of basic_controller.py:82 disappears
        self.agent_in = None
        if getattr(self.args, "stage_agent_inputs", True) and hasattr(mac, "_build_inputs"):
            row = self.env.obs_size
            row += self.env.m if getattr(self.args, "obs_last_action", False) else 0
            row += self.env.n if getattr(self.args, "obs_agent_id", False) else 0
            dtype = th.float32
            # opt-in (args.agent_fc1 = "fp16_split"): the agent's first layer reads the fp16 observation rows themselves
            # (split-precision tensor-core GEMM, modules/agents); the env kernel then stages fp16 rows, padded to a
            # multiple of 8 columns, instead of widening every row to fp32
            if getattr(self.args, "agent_fc1", "fp32") == "fp16_split":
                ok = self.env.kind == "real" and scheme["obs"]["dtype"] == th.float16 and \
                    bool(self.env.lib.sap_real_agent_in_f16_ok(self.env.dims()))
                if not ok:
                    raise ValueError("agent_fc1='fp16_split' needs the real env with its fp16 scheme at a shape the "
                                     "one-CTA-per-env kernel takes (M = N = 10, L = 3, 64 < n <= 128, m <= 128)")
                dtype, row = th.float16, (row + 7) // 8 * 8
            self.agent_in = th.zeros(self.batch_size, self.env.n, row, dtype=dtype, device=self.device)
        # args.overlap_obs_build (default on): build the next observation next to the agent forward (second stream)
        self._overlap = bool(getattr(self.args, "overlap_obs_build", True)) and self.env.kind == "real"
        self._agent_in_bufs, self._side_stream = None, None
        if self._overlap:
            self._side_stream = th.cuda.Stream(device=self.device)
            if self.agent_in is not None:
                self._agent_in_bufs = [self.agent_in, th.zeros_like(self.agent_in)]

    def get_env(self):
        """The reference returns worker 0's env, pickled through the Pipe (parallel_runner.py:246-247, 281-282): a COPY
        with the full env API (``beta_hat``, ``step``, ... - HAAL deep-copies and steps it).  Here: a single-env facade
        built from the same ``env_args``; when they cannot build one (planes adopted with ``set_planes``) the shape /
        scheme record that ``run.py`` reads (run.py:113-131)."""
        if getattr(self, "_env_facade", None) is None:
            self._env_facade = _EnvInfo(self.env)
            ea = dict(self.args.env_args)
            S = ea.get("sat_prox_mat", None)
            try:
                if S is not None and len(S.shape) == 4:
                    S0 = S[self.env_offset if self.global_batch else 0]
                    ea["sat_prox_mat"] = S0.cpu().numpy() if isinstance(S0, th.Tensor) else np.asarray(S0)
                elif isinstance(S, th.Tensor):
                    ea["sat_prox_mat"] = S.cpu().numpy()
                if S is not None and tuple(ea["sat_prox_mat"].shape) == (self.env.n, self.env.m, self.env.T):
                    facade = env_REGISTRY[self.args.env](**ea)
                    facade.obs_size = self.env.obs_size
                    self._env_facade = facade
            except (NotImplementedError, TypeError, KeyError):
                pass
        return self._env_facade

    def get_env_info(self):
        return self.get_env().get_env_info()

    def save_replay(self):
        raise NotImplementedError("the SAP envs have no replay format (the reference's env.save_replay does not exist either)")

    def close_env(self):
        return None

    def attach_replay(self, buffer):
        """Opt-in: roll episodes out directly into ``buffer``'s next ring rows (``ReplayBuffer.view_next``), so the
        following ``buffer.insert_episode_batch(batch)`` is bookkeeping only.  Falls back to a private batch when
        the rows would wrap."""
        self._replay = buffer

    def reset(self, test_mode=False, **reset_kwargs):
        ring = getattr(self, "_replay", None)
        # test episodes are never inserted: they must not be rolled out over sampleable rows of the ring
        view = ring.view_next(self.batch_size) if ring is not None and not test_mode else None
        if view is not None:
            self.batch = view
        elif (self.reuse_batch or getattr(self.args, "use_cuda_graph", False)) and self._batch is not None:
            # (CUDA graphs bake buffer addresses in: a graph runner always rolls out into the same private batch)
            self.batch = self._batch
            for v in self.batch.data.transition_data.values():
                v.zero_()
        else:
            self.batch = self.new_batch()
            self._batch = self.batch
        if self.env.kind == "real":
            self.batch.top_agent_tasks = self.env.top
        self.batch.agent_in = getattr(self, "agent_in", None)   # (= buffer 0 of the double-buffered rows)
        self.batch._agent_in_ids_done = set()   # the MAC rewrites the agent-id columns once per staging buffer and episode
        self.env.reset(self.batch, **reset_kwargs)
        if "beta" in self.lazy:
            # a lazily rebuilt `beta`: the batch rows remember which planes (row, generation) they were rolled out on
            source, generation = self.env.register_planes()
            self.batch.bind_benefit_source(source, generation)
        self.batch.agent_in_t = 0
        self.kernel_launches += self.env.launches_per_step
        self.t = 0

    def _rollout_loop_overlapped(self, test_mode):
        """The T-step loop with the observation build taken off the critical path.  The observation of slot t + 1 depends on
        the benefit window only - not on the actions of step t, apart from M flag columns per agent - so
        ``sap_real_obs_ahead`` builds it on a second stream WHILE the agent network and the selector work on slot t; the
        step itself (``sap_real_step_after_obs``: rewards, counters, flags) then costs microseconds.  The fp32 / fp16
        agent-input rows are double-buffered because the agent is still reading slot t's rows."""
        main = th.cuda.current_stream(self.device)
        side = self._side_stream
        self.mac.init_hidden(batch_size=self.batch_size)
        bufs = self._agent_in_bufs
        order = getattr(self.args, "overlap_submit_order", "agent_first")
        fused = self._fused_select()
        for t in range(self.T):
            nxt = None if bufs is None else bufs[(t + 1) % 2]
            # step t - 1 is done at this point of the main stream: k = t, and nobody reads the rows about to be overwritten
            ready = th.cuda.Event()
            ready.record(main)

            def ahead():
                side.wait_event(ready)
                with th.cuda.stream(side):
                    self.env.obs_ahead(self.batch, agent_in=nxt)

            if order != "agent_first":
                ahead()
            if bufs is not None:
                self.batch.agent_in = bufs[t % 2]
            self.batch.top_agent_tasks = self.env.top   # top-M tasks of slot t (slot t + 1's go to the other buffer)
            if fused:
                # selection happens inside the step launch (sap_rollout_step): only the agent forward precedes it
                q = self.mac.forward(self.batch, t, test_mode=test_mode, action_selection_mode=True)
                if order == "agent_first":
                    ahead()
                main.wait_stream(side)
                sel, actions, keep = self.mac.action_selector.fused_select_args(q, self.t_env, test_mode=test_mode)
                self.env.step_select(sel, actions, self.batch, agent_in=nxt)
                del keep
            else:
                actions = self.mac.select_actions(self.batch, t_ep=t, t_env=self.t_env, test_mode=test_mode)
                if order == "agent_first":
                    # submitted AFTER the agent's GEMMs: those take the SMs first and the observation kernel's CTAs (75 KB
                    # of shared memory each) fill what is left, instead of locking the GEMMs out
                    ahead()
                main.wait_stream(side)
                self.env.step(actions, self.batch, agent_in=nxt)
            self.batch.agent_in_t = t + 1
            if bufs is not None:
                self.batch.agent_in = nxt
            self.t += 1
        self.batch.top_agent_tasks = self.env.top
        self.episode_ctr += 1

    def _fused_select(self):
        """args.fuse_select_step: the classic epsilon-greedy selection runs inside the env step launch (``sap_rollout_step``)
        when MAC, selector and env shape allow it (BasicMAC + ``epsilon_greedy`` + the shipped real-env configuration).
        True (default): with the overlapped schedule, where the step launch is the light ``sap_real_step_after_obs`` kernel;
        "always": also in front of the full step kernel (measured slower there: the Q rows are read at the head of a kernel
        that is already latency-bound); False: never."""
        mode = getattr(self.args, "fuse_select_step", True)
        if not mode or type(self.mac).__name__ != "BasicMAC" or not hasattr(self.mac, "supports_select_and_step"):
            return False
        if mode != "always" and not (self._overlap and self.env.supports_obs_ahead(self.batch)):
            return False
        return self.mac.supports_select_and_step(self.env, self.batch)

    def _rollout_loop(self, test_mode):
        if self._overlap and self.env.supports_obs_ahead(self.batch):
            return self._rollout_loop_overlapped(test_mode)
        self.mac.init_hidden(batch_size=self.batch_size)
        fused = self._fused_select()
        for t in range(self.T):
            if fused:
                # agent forward (torch), then ONE launch: selection + env step (sap_rollout_step)
                self.mac.select_and_step(self.batch, t, self.t_env, self.env, test_mode=test_mode)
            else:
                # agent forward (torch) + selection kernel; obs / avail / beta for slot t were written by the env kernel
                actions = self.mac.select_actions(self.batch, t_ep=t, t_env=self.t_env, test_mode=test_mode)
                # fused env kernel: rewards/actions/terminated at slot t, obs/prev_assigns/filled at slot t+1
                self.env.step(actions, self.batch)
            self.batch.agent_in_t = t + 1
            self.t += 1
        self.episode_ctr += 1

    def _rollout_graph(self, test_mode):
        """Replay the T-step loop as one CUDA graph (``args.use_cuda_graph``).  Every launch reads its step counter,
        episode counter and epsilon from device memory, so one capture serves all later episodes that use the same
        buffers.  The first episode runs eagerly (lazy initialisation), the second is captured."""
        sel = self.mac.action_selector
        if not hasattr(sel, "use_device_epsilon") or not getattr(sel, "graph_capturable", False):
            return False  # e.g. the assignment selectors draw with torch / numpy RNG on the host side
        if not getattr(self.mac, "graph_capturable", False) or getattr(self.mac, "jumpstart_action_selector", None) is not None:
            # JumpstartMAC decides HAA-vs-network per step with host RNG against a host epsilon schedule: a capture
            # would freeze that pattern.  Only MACs that declare themselves capturable (BasicMAC) are replayed.
            return False
        sel.use_device_epsilon(True)
        sel.set_device_epsilon(self.t_env, test_mode, self.device)
        # the captured launches bake in the address of EVERY buffer they touch: all of them go into the key, and the
        # cache entry keeps the tensors alive so that a later allocation can not reuse an address under a stale graph
        view_ptrs = tuple(v.data_ptr() for _, v in sorted(self.batch.data.transition_data.items()))
        extras = (self.env.planes, self.env.plane_stats, self.batch.agent_in, getattr(self.env, "top", None),
                  getattr(self.env, "scratch", None))
        key = view_ptrs + tuple(0 if x is None else x.data_ptr() for x in extras) + (bool(test_mode),)
        graphs = self.__dict__.setdefault("_graphs", {})
        entry = graphs.get(key)
        if entry is None:
            if not self.__dict__.get("_graph_warm", False):
                self._graph_warm = True   # first episode runs eagerly (lazy initialisation inside torch / cuBLAS)
                return False
            if len(graphs) >= self.max_graphs:
                if not self.__dict__.get("_graph_full_warned", False):
                    self._graph_full_warned = True
                    self.logger.console_logger.warning(
                        "CudaVecRunner: %d CUDA graphs cached and the rollout buffers changed again; running eagerly. "
                        "Keep the buffers stable (reuse_episode_batch / attach_replay with batch_size_run dividing the "
                        "ring) or raise runner.max_graphs.", len(graphs))
                return False
            g = th.cuda.CUDAGraph()
            th.cuda.synchronize(self.device)
            t_host = self.env.t_host
            agent = getattr(self.mac, "agent", None)
            a0 = getattr(agent, "kernel_launches", 0)
            with th.cuda.graph(g):
                self._rollout_loop(test_mode)
            self.env.t_host = t_host  # capture only recorded the launches
            self.t = 0
            # this repo's kernels inside the agent forward (bias / ReLU epilogues) per captured episode: the host counter
            # does not advance on a replay, so the replay adds them
            entry = (g, self.batch, extras, getattr(agent, "kernel_launches", 0) - a0)
            if agent is not None and entry[3]:
                agent.__dict__["kernel_launches"] = a0
            graphs[key] = entry
        g = entry[0]
        g.replay()
        if entry[3]:
            self.mac.agent.__dict__["kernel_launches"] = getattr(self.mac.agent, "kernel_launches", 0) + entry[3]
        self.env.t_host += self.T
        self.t = self.T
        self.batch.agent_in_t = self.T
        return True

    @th.no_grad()
    def run(self, test_mode=False, **reset_kwargs):
        self.reset(test_mode=test_mode, **reset_kwargs)
        if not (getattr(self.args, "use_cuda_graph", False) and self._rollout_graph(test_mode)):
            self._rollout_loop(test_mode)
        if self.compat_quirks:
            self._apply_parallel_runner_quirks(test_mode)
        per_step = 1 + self.env.launches_per_step  # selector + env kernel(s) per timestep
        if self._overlap and self.env.supports_obs_ahead(self.batch):
            per_step += 1                           # observation kernel + step kernel instead of one fused launch
        if self._fused_select():
            per_step -= 1                           # the selector runs inside the step launch
        self.kernel_launches += per_step * self.T
        self.last_episode_returns = self.env.ep_return.clone()
        self._finish_run(test_mode)
        return self.batch

    def _apply_parallel_runner_quirks(self, test_mode):
        """What the reference ParallelRunner stores beyond EpisodeRunner semantics (golden
        tests/golden/runner_parallel.npz), reproduced only under ``compat_parallel_runner_quirks``:

        * parallel_runner.py:131-139: the loop selects actions once more at t = T (for envs it has not yet noticed to be
          finished) and stores them - plus their one-hot - in slot T before it breaks;
        * parallel_runner.py:181-187: ``terminated`` is assigned from the truthiness of the reward LIST collected so far
          in this timestep, not from the env: False for the first env of the batch (the list is still empty), True for
          every other env, at every step t < T."""
        td = self.batch.data.transition_data
        B, n, T = self.batch_size, self.env.n, self.T
        actions = self.mac.select_actions(self.batch, t_ep=T, t_env=self.t_env, test_mode=test_mode)
        self.kernel_launches += 1
        if "actions" in td and td["actions"].shape[-1] == 1:
            td["actions"][:, T] = actions.view(B, n, 1).to(td["actions"].dtype)
            if "actions_onehot" in td:
                oh = td["actions_onehot"][:, T]
                oh.zero_()
                oh.scatter_(2, actions.view(B, n, 1).long(), 1)
        if "terminated" in td:
            td["terminated"][:, :T] = True
            if self.env_offset == 0:
                td["terminated"][0, :T] = False

    # ------------------------------------------------------------------ statistics (A.6 of SURVEY.md)
    def _finish_run(self, test_mode):
        ret = self.last_episode_returns
        mom = th.stack([ret.sum(), (ret * ret).sum(),
                        th.tensor(float(self.batch_size), dtype=th.float64, device=ret.device),
                        th.tensor(float(self.batch_size * self.T), dtype=th.float64, device=ret.device)])
        world = 1
        if th.distributed.is_available() and th.distributed.is_initialized():
            th.distributed.all_reduce(mom)  # the only collective of a rollout
            world = th.distributed.get_world_size()
        s, s2, n_ep, ep_len = mom.tolist()  # one host read per run()
        cur_stats = self.test_stats if test_mode else self.train_stats
        cur_returns = self.test_returns if test_mode else self.train_returns
        log_prefix = "test_" if test_mode else ""
        cur_stats["n_episodes"] = int(n_ep) + cur_stats.get("n_episodes", 0)
        cur_stats["ep_length"] = int(ep_len) + cur_stats.get("ep_length", 0)
        cur_returns.add(n_ep, s, s2)
        if not test_mode:
            self.t_env += int(ep_len)

        n_test_runs = max(1, self.args.test_nepisode // (self.batch_size * world)) * self.batch_size * world
        if test_mode and (cur_returns.count == n_test_runs):
            self._log(cur_returns, cur_stats, log_prefix)
        elif self.t_env - self.log_train_stats_t >= self.args.runner_log_interval:
            self._log(cur_returns, cur_stats, log_prefix)
            if hasattr(self.mac.action_selector, "epsilon"):
                self.logger.log_stat("epsilon", self.mac.action_selector.epsilon, self.t_env)
            self.log_train_stats_t = self.t_env
            self.logger.log_stat("steps", self.t_env, self.t_env)

    def _log(self, returns, stats, prefix):
        self.logger.log_stat(prefix + "return_mean", returns.mean(), self.t_env)
        self.logger.log_stat(prefix + "return_std", returns.std(), self.t_env)
        returns.clear()
        for k, v in stats.items():
            if k != "n_episodes":
                self.logger.log_stat(prefix + k + "_mean", v / stats["n_episodes"], self.t_env)
        stats.clear()


class _Moments:
    """Sufficient statistics of the episode returns since the last log (np.mean / np.std, population)."""

    def __init__(self):
        self.clear()

    def clear(self):
        self.count, self.s, self.s2 = 0, 0.0, 0.0

    def add(self, n, s, s2):
        self.count += int(n)
        self.s += s
        self.s2 += s2

    def __len__(self):
        return self.count

    def mean(self):
        return self.s / self.count if self.count else float("nan")

    def std(self):
        if not self.count:
            return float("nan")
        mu = self.s / self.count
        return math.sqrt(max(self.s2 / self.count - mu * mu, 0.0))


class ParallelRunner(CudaVecRunner):
    """REGISTRY["parallel"]: ``batch_size_run`` envs per launch (parallel_runner.py:12-243)."""


class EpisodeRunner(CudaVecRunner):
    """REGISTRY["episode"]: one env (episode_runner.py:8-137)."""

    def __init__(self, args, logger):
        assert args.batch_size_run == 1, "EpisodeRunner only supports batch size 1"
        super().__init__(args, logger)
        self.log_train_stats_t = -1000000
