"""Throughput of the rollout hot path (agent-steps/s) on N B200s, with the step kernel's HBM roofline and the
reference's own CPU runner timed beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4|c1|tiny] [--impl reference]

A "step" = one full rollout of the hot path over this rank's batch of synthetic environments through the
reference-facing API: ``runner.run()`` (T timesteps of: torch agent forward -> selection kernel -> env step / obs /
buffer-write kernel; the T-step loop is replayed as one CUDA graph unless --no-graph) followed by
``ReplayBuffer.insert_episode_batch``.  Prints ONE JSON line on rank 0.

value   : whole-job agent-steps/s with the benefit tensors already resident in HBM, default configuration (fp32 agent,
          beta / avail_actions / actions_onehot lazy).  Weak scaling: every rank owns the workload's envs.
e2e     : the same call with host inputs and outputs: agent parameters H2D from pinned memory every step (what the
          learner returns between rollouts), per-env returns / actions / rewards D2H; the env's benefit tensor is constant
          across episodes as in the reference, its one-time upload is reported under e2e.setup.
          (--e2e-fresh-benefits adds the round-1 definition: all benefits re-uploaded every episode.)
roofline: the env kernel; achieved = SURVEY.md 8(d) bytes per launch / mean launch duration from CUDA events around every
          launch (one extra eager episode when the timed region replayed graphs); the bytes this design really moves
          (fp16 obs + agent-input staging) are reported beside it as design_bytes / frac_design_bytes.
variants: the same workload with the opt-in split-precision first layer (args.agent_fc1 = "fp16_split") and with every
          buffer field materialised (eager), each a short run.
strong_scaling (N > 1): the SAME env count split over the ranks (batch_size_run_is_global): the north star's
          "4096 envs sharded across 2/4/8 GPUs".
cpu_baseline / --impl reference: the UNMODIFIED reference (oracle/_ref, see oracle/make_ref.py) - ParallelRunner with one
          env process per host core + BasicMAC + RNNAgent + epsilon-greedy - on a bounded sample; the numpy oracle port is
          timed too and reported as cpu_baseline.port.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: per-GPU envs, agents, tasks, T, L, M, N, env
    "c3": dict(B=4096, n=100, m=100, T=100, L=3, M=10, N=10, env="real", desc="4096 envs x 100 agents x 100 tasks per GPU"),
    "c2": dict(B=1024, n=50, m=50, T=100, L=3, M=10, N=10, env="real", desc="1024 envs x 50 agents x 50 tasks per GPU"),
    "c4": dict(B=64, n=324, m=450, T=100, L=3, M=10, N=10, env="real", desc="64 envs x 324 agents x 450 tasks per GPU"),
    "c1": dict(B=1, n=10, m=10, T=100, L=3, M=0, N=0, env="mock", desc="mock env 10 agents x 10 tasks, one env"),
    "c3mock": dict(B=4096, n=100, m=100, T=100, L=3, M=0, N=0, env="mock", desc="mock env 4096 x 100 x 100 per GPU"),
    "tiny": dict(B=64, n=20, m=30, T=10, L=3, M=6, N=4, env="real", desc="tiny functional check"),
}


def obs_size(w):
    if w["env"] == "mock":
        return (w["L"] + 1) * w["m"]
    return w["M"] * w["L"] + w["N"] * w["M"] * w["L"] + w["N"] * (w["M"] // 2) * w["L"] + w["M"]


def algorithmic_bytes_per_env_step(w, e_obs, e_agent_in=4):
    """DESIGN.md section 4 / SURVEY.md 8(d): window read + obs write (buffer dtype) + the fp32 copy of the obs the
    agent network consumes (written by the same kernel) + per-agent scalars."""
    n, m, L = w["n"], w["m"], w["L"]
    return n * m * L * 4 + n * obs_size(w) * (e_obs + e_agent_in) + n * 16 + 16


# ----------------------------------------------------------------------------------------------- CPU port
def _agent_numpy_forward(weights, x):
    import numpy as np

    h = np.maximum(x @ weights["fc1.weight"].T + weights["fc1.bias"], 0)
    h = np.maximum(h @ weights["rnn.weight"].T + weights["rnn.bias"], 0)
    return h @ weights["fc2.weight"].T + weights["fc2.bias"]


def _cpu_worker(job):
    """One host core: the oracle env + the same fc agent + oracle selector + buffer writes for `Bc` envs."""
    import numpy as np
    import torch as th

    from oracle import cpu_oracle as O

    th.set_num_threads(1)
    w, Bc, t_steps, seed, weights = job
    rng = np.random.default_rng(seed)
    n, m, L = w["n"], w["m"], w["L"]
    T = min(w["T"], t_steps + L)  # only the planes the sampled timesteps read (per-step work is unchanged)
    S = rng.random((Bc, n, m, T), dtype=np.float32).astype(np.float64)
    real = w["env"] == "real"
    state = O.RealState(S, L, w["M"], w["N"], 0.5) if real else O.MockState(S, L, 0.5)
    real_dt = np.float16 if real else np.float32
    osz = obs_size(w)
    buf_obs = np.zeros((Bc, t_steps + 1, n, osz), dtype=real_dt)
    buf_rew = np.zeros((Bc, t_steps + 1, n), dtype=real_dt)
    buf_act = np.zeros((Bc, t_steps + 1, n), dtype=np.int16 if real else np.int64)
    tw = {k: th.tensor(v) for k, v in weights.items()}
    t0 = time.perf_counter()
    if real:
        state.reset()
    else:
        state.reset(np.stack([rng.permutation(m)[:n] for _ in range(Bc)]))
    avail = np.ones((Bc, n, m), dtype=bool)
    for t in range(t_steps):
        pre = state.pretransition()
        buf_obs[:, t] = pre["obs"]
        x = th.from_numpy(buf_obs[:, t].astype(np.float32).reshape(Bc * n, -1))
        h = th.relu(th.nn.functional.linear(x, tw["fc1.weight"], tw["fc1.bias"]))
        h = th.relu(th.nn.functional.linear(h, tw["rnn.weight"], tw["rnn.bias"]))
        q = th.nn.functional.linear(h, tw["fc2.weight"], tw["fc2.bias"]).numpy().reshape(Bc, n, m)
        a = O.select_epsilon_greedy(q, avail, 0.5, rng.random((Bc, n), dtype=np.float32), rng.random((Bc, n), dtype=np.float32))
        r, done = state.step(a)
        buf_rew[:, t], buf_act[:, t] = r, a
    buf_obs[:, t_steps] = state.pretransition()["obs"]
    return time.perf_counter() - t0, Bc * n * t_steps


def cpu_port_throughput(w, weights, target_seconds=12.0, procs=None):
    """agent-steps/s of the CPU port on `procs` host cores over a bounded sample of the workload."""
    import multiprocessing as mp

    import psutil

    procs = procs or len(os.sched_getaffinity(0)) or 1
    # calibrate on one core, one env, two steps
    dt, steps = _cpu_worker((w, 1, 2, 0, weights))
    per_env_step = dt / 2
    t_steps = min(w["T"], 10)
    Bc = max(1, min(w["B"], int(target_seconds / max(per_env_step * t_steps, 1e-6))))
    per_env_bytes = w["n"] * w["m"] * (t_steps + w["L"]) * 8 * 4 + w["n"] * obs_size(w) * (t_steps + 1) * 4
    Bc = max(1, min(Bc, int(0.25 * psutil.virtual_memory().available / procs / per_env_bytes)))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_worker, [(w, Bc, t_steps, 100 + i, weights) for i in range(procs)])
    wall = max(r[0] for r in res)
    total = sum(r[1] for r in res)
    sample = f"{procs} procs x {Bc} envs x {t_steps} timesteps of {w['n']}x{w['m']} ({w['env']} env, oracle + fc agent, numpy)"
    return total / wall, procs, sample, time.perf_counter() - t0


def default_agent_weights(w, hidden=64, seed=0):
    import numpy as np

    rng = np.random.default_rng(seed)
    d_in, d_out = obs_size(w), w["m"]

    def lin(o, i):
        bound = 1.0 / np.sqrt(i)
        return (rng.uniform(-bound, bound, (o, i)).astype(np.float32), rng.uniform(-bound, bound, o).astype(np.float32))

    f1, r1, f2 = lin(hidden, d_in), lin(hidden, hidden), lin(d_out, hidden)
    return {"fc1.weight": f1[0], "fc1.bias": f1[1], "rnn.weight": r1[0], "rnn.bias": r1[1], "fc2.weight": f2[0],
            "fc2.bias": f2[1]}


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ----------------------------------------------------------------------------------------------- GPU arm
def bytes_8d(w):
    """SURVEY.md 8(d) / BASELINE.md section 4: window read + obs write at 4 bytes + per-agent scalars, per env-step."""
    n, m, L = w["n"], w["m"], w["L"]
    return n * m * L * 4 + n * obs_size(w) * 4 + n * 16 + 16


def build_runner(w, rank_seed, planes, use_graph=True, selector="epsilon_greedy", lazy=("beta", "avail_actions", "actions_onehot"),
                 agent_fc1="fp32", global_batch=None, overlap=True, fuse_select=True):
    import torch as th

    from marl_sap_b200.components.episode_buffer import ReplayBuffer
    from marl_sap_b200.controllers import REGISTRY as mac_REGISTRY
    from marl_sap_b200.runners import REGISTRY as r_REGISTRY
    from marl_sap_b200.utils.logging import Logger

    real = w["env"] == "real"
    # the runner adopts the device-resident planes directly; a 1-env placeholder satisfies the env_args contract
    placeholder = th.zeros(w["n"], w["m"], w["T"])
    if real:
        env_args = dict(num_planes=1, num_sats_per_plane=w["n"], m=w["m"], T=w["T"], N=w["N"], M=w["M"], L=w["L"],
                        lambda_=0.5, sat_prox_mat=placeholder, graphs=1)
        env_name = "real_constellation_env"
    else:
        env_args = dict(n=w["n"], m=w["m"], T=w["T"], L=w["L"], lambda_=0.5, sat_prox_mat=placeholder)
        env_name = "mock_constellation_env"
    args = SimpleNamespace(env=env_name, env_args=env_args, batch_size_run=global_batch or planes.shape[0], device="cuda",
                           runner="parallel", batch_size_run_is_global=global_batch is not None,
                           mac="basic_mac", action_selector=selector, epsilon_start=0.5, epsilon_finish=0.5,
                           epsilon_anneal_time=1, evaluation_epsilon=0.0, agent="rnn", hidden_dim=64, use_rnn=False,
                           obs_agent_id=False, obs_last_action=False, agent_output_type="q", test_nepisode=planes.shape[0],
                           runner_log_interval=10 ** 12, seed=rank_seed, use_mps_action_selection=True,
                           lazy_buffer_fields=tuple(lazy), reuse_episode_batch=True, use_cuda_graph=bool(use_graph),
                           agent_fc1=agent_fc1, overlap_obs_build=overlap, fuse_select_step=fuse_select,
                           overlap_submit_order=os.environ.get("SAP_OVERLAP_ORDER", "agent_first"))
    logger = Logger()
    runner = r_REGISTRY["parallel"](args=args, logger=logger)
    assert runner.batch_size == planes.shape[0], (runner.batch_size, planes.shape)
    runner.env.set_planes(planes, shared=False)
    env = runner.get_env()
    args.n, args.m, args.T = env.n, env.m, env.T
    groups = {"agents": args.n}
    buffer = ReplayBuffer(env.scheme, groups, planes.shape[0], env.T + 1, preprocess=env.preprocess, device="cuda",
                          lazy=args.lazy_buffer_fields)
    th.manual_seed(0)
    mac = mac_REGISTRY["basic_mac"](buffer.scheme, groups, args)
    weights = default_agent_weights(w)
    mac.agent.load_state_dict({k: th.tensor(v) for k, v in weights.items()})
    mac.cuda()
    runner.setup(scheme=env.scheme, groups=groups, preprocess=env.preprocess, mac=mac)
    runner.attach_replay(buffer)  # episodes are rolled out in the replay ring's own rows (no insert-time copy)
    return runner, buffer, weights


def workload_config(opts, w, world):
    """The `config` object of the JSON line: the workload only, identical in both arms (implementation switches of the GPU
    arm go to `impl_config`)."""
    planes_gib = w["B"] * w["T"] * w["n"] * w["m"] * 4 / 2 ** 30
    return {"workload": f"{opts.workload}: {w['desc']}, T={w['T']}, L={w['L']}, M={w['M']}, N={w['N']}, {w['env']} env, "
                        f"{opts.selector} + fc agent(hidden 64)",
            "envs_per_gpu": w["B"], "agents": w["n"], "tasks": w["m"], "T": w["T"],
            "step": "runner.run() + ReplayBuffer.insert_episode_batch",
            "inputs": (f"benefit planes {planes_gib:.1f} GiB per GPU (> 126 MB L2), distinct per env" if planes_gib > 0.2 else
                       f"benefit planes {planes_gib * 1024:.2f} MiB per GPU: smaller than L2 and not flushed (a latency-bound "
                       "workload; the L2-independent lines are c2 / c3 / c4)"),
            "env_arithmetic": "f64 sums/rewards on f32 benefits; obs/rewards stored in the scheme dtype",
            "parallelism": f"envs block-partitioned, {world} rank(s), no data-path collective"}


class Rollout:
    """runner + buffer with per-launch timing of the env kernel and the timed-step helper shared by every measurement."""

    def __init__(self, th, dist, world, runner, buffer):
        self.th, self.dist, self.world, self.runner, self.buffer = th, dist, world, runner, buffer
        self.ev_pairs, self.timing = [], False
        orig = runner.env.step

        def timed_step(actions, batch, **kw):
            if self.timing:
                a, b = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
                a.record()
                r = orig(actions, batch, **kw)
                b.record()
                self.ev_pairs.append((a, b))
                return r
            return orig(actions, batch, **kw)

        runner.env.step = timed_step

    def step(self):
        with self.th.no_grad():
            batch = self.runner.run(test_mode=False)
        self.buffer.insert_episode_batch(batch)
        return batch

    def launches(self):
        return self.runner.kernel_launches + self.buffer.kernel_launches + getattr(self.runner.mac.agent, "kernel_launches", 0)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def measure(self, steps, warmup, graph, sampler=None):
        """(ms for `steps` episodes [max over ranks], mean env-kernel launch ms [max over ranks], launches, wall s)."""
        th = self.th
        for _ in range(warmup):
            self.step()
        l0 = self.launches()
        self.barrier()
        th.cuda.synchronize()
        if sampler is not None:
            sampler.start()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        self.timing, self.ev_pairs = not graph, []
        wall0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            self.step()
        e1.record()
        th.cuda.synchronize()
        wall = time.perf_counter() - wall0
        self.barrier()
        self.timing = False
        launches = self.launches() - l0
        overlapped = getattr(self.runner, "_overlap", False)
        fused_sel = getattr(self.runner.args, "fuse_select_step", True)
        if graph or overlapped or not self.ev_pairs:
            # the timed region replayed captured CUDA graphs, built the observations on a second stream next to the agent's
            # GEMMs and / or selected inside the step launch (no clean per-launch events of the env step possible): the
            # env kernel's own launch duration comes from one extra episode right after it, launched eagerly with the
            # full step (sap_real_step: rewards + observation build, no selection) alone on the stream
            self.runner.args.use_cuda_graph = False
            self.runner.args.fuse_select_step = False
            self.runner._overlap = False
            self.timing, self.ev_pairs = True, []
            self.step()
            th.cuda.synchronize()
            self.timing = False
            self.runner.args.use_cuda_graph = graph
            self.runner.args.fuse_select_step = fused_sel
            self.runner._overlap = overlapped
        ms = e0.elapsed_time(e1)
        kern = sum(a.elapsed_time(b) for a, b in self.ev_pairs) / max(len(self.ev_pairs), 1)
        t = th.tensor([ms, kern], dtype=th.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        ms, kern = t.tolist()
        return ms, kern, launches, wall


def gpu_arm(opts, w):
    import gc

    import torch as th
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    th.cuda.set_device(local_rank)
    dev = th.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, n, m, T = w["B"], w["n"], w["m"], w["T"]
    graph = not opts.no_graph and opts.selector == "epsilon_greedy"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    e_obs = 2 if w["env"] == "real" else 4

    def roof(kern_ms, envs, e_ain):
        """achieved GB/s by the SURVEY 8(d) numerator and by what this configuration really has to move"""
        a8 = bytes_8d(w) * envs / (kern_ms * 1e-3) / 1e9
        ad = algorithmic_bytes_per_env_step(w, e_obs, e_ain) * envs / (kern_ms * 1e-3) / 1e9
        return a8, ad

    # synthetic benefits, U(0,1), distinct per env and per rank, generated straight in the device layout
    g = th.Generator(device=dev).manual_seed(1234 + rank)
    planes = th.rand(B, T, n, m, device=dev, generator=g)

    # ------------------------------------------------------------------ headline: default configuration, weak scaling
    runner, buffer, weights = build_runner(w, 1, planes, graph, opts.selector, agent_fc1=opts.agent_fc1)
    ro = Rollout(th, dist, world, runner, buffer)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, kern_ms, launches, wall = ro.measure(opts.steps, opts.warmup, graph, sampler)
    clocks = sampler.stop() if rank == 0 else None
    agent_steps = world * B * n * T * opts.steps
    value = agent_steps / (ms * 1e-3)
    e_ain = 0 if runner.agent_in is None else runner.agent_in.element_size()
    multi_cta = getattr(runner.env, "launches_per_step", 1) == 4
    runner_overlap = SimpleNamespace(value=bool(getattr(runner, "_overlap", False)) and runner.env.supports_obs_ahead(runner.batch),
                                     fused_select=bool(runner._fused_select()))

    # ------------------------------------------------------------------ e2e: host inputs in, host results out
    e2e, e2e_fresh = None, None
    if not opts.no_e2e:
        e2e = e2e_leg(opts, w, ro, planes, weights, dev, world)
        if opts.e2e_fresh_benefits:
            e2e_fresh = e2e_fresh_benefits_leg(opts, w, ro, planes, dev, world)
    del ro, runner, buffer
    gc.collect()
    th.cuda.empty_cache()

    # ------------------------------------------------------------------ variants of the same workload (short runs)
    variants = {}

    def variant(name, note, **kw):
        try:
            r, bf, _ = build_runner(w, 1, planes, graph, opts.selector, **kw)
            x = Rollout(th, dist, world, r, bf)
            vms, vk, _, _ = x.measure(max(2, min(opts.steps, 3)), 2, graph)
            ea = 0 if r.agent_in is None else r.agent_in.element_size()
            a8, ad = roof(vk, B, ea)
            variants[name] = {"value": world * B * n * T * max(2, min(opts.steps, 3)) / (vms * 1e-3), "unit": "agent-steps/s",
                              "ms_per_step": vms / max(2, min(opts.steps, 3)), "env_kernel_ms": vk,
                              "roofline_frac": a8 / peak, "roofline_frac_design_bytes": ad / peak, "note": note}
            del x, r, bf
        except Exception as e:  # a variant must never take the headline down (e.g. out of memory for the eager buffer)
            variants[name] = {"value": None, "error": f"{type(e).__name__}: {e}"[:300]}
        gc.collect()
        th.cuda.empty_cache()

    if not opts.no_variants and w["env"] == "real" and not multi_cta and opts.selector == "epsilon_greedy":
        other = "fp32" if opts.agent_fc1 == "fp16_split" else "fp16_split"
        variant("agent_fc1_" + other,
                "opt-in args.agent_fc1='fp16_split': the env kernel stages fp16 rows (padded to 496 columns) instead of fp32 "
                "rows and fc1 is ONE fp16 tensor-core GEMM against [W0|W1|W2] with fp32 accumulation + a fold/bias/ReLU "
                "kernel (max rel. error vs float64 9.5e-7; the fp32 sgemm: 1.4e-6)" if other == "fp16_split" else
                "default fp32 agent: fp32 staging rows, torch/cuBLAS sgemm", agent_fc1=other)
        variant("separate_selector", "args.fuse_select_step=False: sap_select_epsilon_greedy and the env step as two launches "
                "(default: sap_rollout_step selects inside the step launch of the overlapped schedule)", agent_fc1=opts.agent_fc1, fuse_select=False)
        variant("fused_step", "args.overlap_obs_build=False: the observation is built inside the step kernel, after the "
                "selection (one stream; the round-1 schedule)", agent_fc1=opts.agent_fc1, overlap=False)
        free, _ = th.cuda.mem_get_info(dev)
        eager_bytes = B * (T + 1) * (n * obs_size(w) * 2 + n * m * w["L"] * 2 + n * m * 1 + n * m * 2 + 64 * n)
        if eager_bytes * 1.1 < free:
            variant("eager_buffer", "every scheme field materialised like the reference's EpisodeBatch (beta fp16 [n,m,L], "
                    "avail_actions bool [n,m], actions_onehot int16 [n,m] written by the env kernel each step)", lazy=(),
                    agent_fc1=opts.agent_fc1)
        else:
            variants["eager_buffer"] = {"value": None, "skipped": f"needs {eager_bytes >> 30} GiB, {free >> 30} GiB free"}

    # ------------------------------------------------------------------ strong scaling: the SAME B envs split over the ranks
    strong = None
    if world > 1 and not opts.no_strong:
        if B % world == 0:
            Bl = B // world
            lp = planes[:Bl]
            r, bf, _ = build_runner(w, 1, lp, graph, opts.selector, agent_fc1=opts.agent_fc1, global_batch=B)
            x = Rollout(th, dist, world, r, bf)
            sms, sk, _, _ = x.measure(opts.steps, opts.warmup, graph)
            a8, _ = roof(sk, Bl, 0)
            ctas = Bl * (1 if not multi_cta else 0)
            strong = {"value": B * n * T * opts.steps / (sms * 1e-3), "unit": "agent-steps/s", "total_envs": B, "envs_per_gpu": Bl,
                      "ms_per_step": sms / opts.steps, "us_per_timestep": 1e3 * sms / opts.steps / T, "env_kernel_ms": sk,
                      "env_kernel_frac_8d": a8 / peak,
                      "efficiency_vs_weak": (B * n * T * opts.steps / (sms * 1e-3)) / value,
                      "limiter": (f"kernel time, not launches: {Bl} one-CTA envs on {148 * 3} resident CTA slots = "
                                  f"{Bl / (148 * 3):.2f} waves, so the observation kernel needs {math.ceil(Bl / (148 * 3))} "
                                  f"full env latencies ({sk * 1e3:.0f} us per launch measured alone) and the agent's SIMT sgemms on "
                                  f"{Bl * n} rows take most of the rest of the {1e3 * sms / opts.steps / T:.0f} us timestep; the two "
                                  "take turns on the SMs (an observation CTA needs 75 KB of shared memory and 20 K registers)")
                      if ctas else "multi-CTA path",
                      "note": "batch_size_run_is_global=True: utils.dist.env_partition block-partitions the envs, no data-path "
                              "collective; CUDA-graph replay of the T-step loop" if graph else "eager launches"}
            del x, r, bf
        else:
            strong = {"value": None, "skipped": f"{B} envs do not divide over {world} ranks"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    a8, ad = roof(kern_ms, B, e_ain)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(opts.workload)
    except Exception:
        pass
    cpu = None
    if not opts.no_cpu and world == 1 and opts.selector == "epsilon_greedy":
        cpu = cpu_baseline(w, weights, opts.cpu_seconds)
    if w["env"] != "real":
        kernel_name = "sap_mock_kernel"
    elif multi_cta:  # one env over many CTAs: the four launches of one env step
        kernel_name = "sap_real_large_{prep,keys,lists,main} (4 launches per env step, timed together)"
    else:
        kernel_name = "sap_real_fast2_kernel" if (w["M"], w["N"], w["L"]) == (10, 10, 3) and 64 < n <= 128 and m <= 128 else "sap_real_fast_kernel"
    line = {
        "metric": "agent_steps_per_sec", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": opts.steps,
        "warmup": opts.warmup, "ms_per_step": ms / opts.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(opts, w, world),
        "impl_config": {"cuda_graph": bool(graph), "agent_fc1": opts.agent_fc1,
                        "overlap_obs_build": bool(getattr(runner_overlap, "value", False)),
                        "fuse_select_step": bool(getattr(runner_overlap, "fused_select", False)),
                        "buffer_fields": "obs/actions/rewards/terminated/filled/prev_assigns eager; beta/avail/onehot lazy "
                                         "(rebuilt on access from the episode's planes); episodes rolled out in place in the "
                                         "replay ring; variants.eager_buffer has every field materialised"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": a8, "peak": peak, "unit": "GB/s", "frac": a8 / peak,
                     "traffic": traffic, "kernel": kernel_name,
                     "algorithmic_bytes": "SURVEY.md 8(d): window read n*m*L*4 + obs write n*obs*4 + n*16+16 per env-step",
                     "algorithmic_bytes_per_launch": bytes_8d(w) * B, "avg_launch_ms": kern_ms, "peak_source": peak_src,
                     "kernel_share_of_step": kern_ms * T / (ms / opts.steps),
                     "design_bytes_per_launch": algorithmic_bytes_per_env_step(w, e_obs, e_ain) * B,
                     "design_bytes": f"window n*m*L*4 + obs n*obs*{e_obs} + agent-input staging n*obs*{e_ain} + n*16+16",
                     "achieved_design_bytes": ad, "frac_design_bytes": ad / peak},
        "cpu_baseline": cpu, "wall_s": wall, "variants": variants or None, "strong_scaling": strong,
        "e2e_fresh_benefits": e2e_fresh,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _d2h_buffers(th, runner, B, T, n):
    ret_host = th.empty(B, dtype=th.float64, pin_memory=True)
    act_host = th.empty(B, T + 1, n, 1, dtype=runner.env.scheme["actions"]["dtype"], pin_memory=True)
    rew_host = th.empty(B, T + 1, n, dtype=runner.env.scheme["rewards"]["dtype"], pin_memory=True)
    nbytes = ret_host.numel() * 8 + act_host.numel() * act_host.element_size() + rew_host.numel() * rew_host.element_size()
    return ret_host, act_host, rew_host, nbytes


def e2e_leg(opts, w, ro, planes, weights, dev, world):
    """The same step through the public API with HOST inputs and outputs.

    The benefit tensor is a constant of the env, exactly as in the reference (``sat_prox_mat`` is a constructor argument
    that every episode re-reads, real_constellation_env.py:47-60, 121-127): it is uploaded ONCE, from pinned host memory in
    the reference's [B,n,m,T] layout through ``sap_benefit_upload_host`` (timed and reported as ``setup``, outside the
    per-step region, like the reference's env construction).  What a step receives from the host every episode is what
    the learner hands back between rollouts (run.py:262-283; the reference copies it into its selector agent,
    basic_controller.py:69-75): the agent's parameters, uploaded from pinned memory.  What goes back to the host every
    step: per-env returns, the joint actions and the per-agent rewards of the whole episode."""
    import torch as th
    import torch.distributed as dist

    from marl_sap_b200 import _lib

    runner, buffer = ro.runner, ro.buffer
    B, n, m, T = w["B"], w["n"], w["m"], w["T"]
    lib = _lib.load()
    # ---- one-time benefit upload through the public host path (bounded pinned pool, cycled)
    per_env = n * m * T * 4
    chunk = max(1, min(B, (256 << 20) // per_env))
    host_envs = min(B, max(chunk, ((2 << 30) // per_env) // chunk * chunk))
    setup = None
    try:
        host = th.empty(host_envs, n, m, T, dtype=th.float32, pin_memory=True)
        for b0 in range(0, host_envs, chunk):
            b1 = min(host_envs, b0 + chunk)
            host[b0:b1].copy_(th.rand(b1 - b0, n, m, T, device=dev))
        staging = [th.empty(chunk * n * m * T, dtype=th.float32, device=dev) for _ in range(2)]
        th.cuda.synchronize()
        s0, s1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        s0.record()
        stream = _lib.stream_ptr(dev)
        for i, b0 in enumerate(range(0, B, chunk)):
            b1 = min(B, b0 + chunk)
            h0 = b0 % host_envs
            _lib.check(lib.sap_benefit_upload_host(host[h0:h0 + (b1 - b0)].data_ptr(), staging[i % 2].data_ptr(),
                                                   planes[b0:b1].data_ptr(), b1 - b0, n, m, T, stream), "sap_benefit_upload_host")
        runner.env.set_planes(planes)
        s1.record()
        th.cuda.synchronize()
        setup = {"benefit_h2d_bytes": B * per_env, "benefit_upload_ms": s0.elapsed_time(s1),
                 "note": "once per change of sat_prox_mat, not per episode"}
        del host, staging
    except RuntimeError as e:
        setup = {"skipped": str(e)[:200]}
    # ---- per-step host traffic
    params = [p for p in runner.mac.agent.parameters()]
    flat_host = th.cat([th.tensor(weights[k]).reshape(-1) for k in ("fc1.weight", "fc1.bias", "rnn.weight", "rnn.bias",
                                                                      "fc2.weight", "fc2.bias")]).pin_memory()
    flat_dev = th.empty_like(flat_host, device=dev)
    names = [k for k, _ in runner.mac.agent.named_parameters()]
    assert names == ["fc1.weight", "fc1.bias", "rnn.weight", "rnn.bias", "fc2.weight", "fc2.bias"], names
    ret_host, act_host, rew_host, bytes_out = _d2h_buffers(th, runner, B, T, n)
    bytes_in = flat_host.numel() * 4

    def e2e_step():
        flat_dev.copy_(flat_host, non_blocking=True)          # H2D: refreshed agent parameters
        off = 0
        with th.no_grad():
            for p in params:
                p.copy_(flat_dev[off:off + p.numel()].view_as(p))
                off += p.numel()
        batch = ro.step()
        ret_host.copy_(runner.last_episode_returns, non_blocking=True)
        act_host.copy_(batch["actions"], non_blocking=True)
        rew_host.copy_(batch["rewards"], non_blocking=True)

    steps = max(2, min(opts.steps, opts.e2e_steps))
    e2e_step()
    e2e_step()
    th.cuda.synchronize()
    ro.barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        e2e_step()
    e1.record()
    th.cuda.synchronize()
    ms = th.tensor([e0.elapsed_time(e1)], dtype=th.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    val = world * B * n * T * steps / (ms.item() * 1e-3)
    return {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out,
            "steps": steps, "ms_per_step": ms.item() / steps, "setup": setup,
            "note": "per step: agent parameters H2D from pinned memory (what the learner returns between rollouts), "
                    "runner.run() + ReplayBuffer.insert_episode_batch, then returns / actions / rewards D2H into pinned "
                    "memory; the env's benefit tensor is constant across episodes as in the reference and is uploaded once "
                    "(setup); e2e_fresh_benefits re-uploads all benefits every episode (the round-1 definition)"}


def e2e_fresh_benefits_leg(opts, w, ro, planes, dev, world):
    """Round-1 definition, kept for continuity: EVERY episode the benefit tensors of all envs are uploaded from pinned
    memory in the reference layout [B,n,m,T] (H2D + re-layout, double-buffered on a copy stream so that the upload of
    episode e+1 overlaps the rollout of episode e).  PCIe / host-memory bound; the reference never does this."""
    import psutil
    import torch as th
    import torch.distributed as dist

    from marl_sap_b200 import _lib

    runner, buffer = ro.runner, ro.buffer
    B, n, m, T = w["B"], w["n"], w["m"], w["T"]
    per_env = n * m * T * 4
    bytes_in = B * per_env
    chunk = max(1, min(B, (256 << 20) // per_env))           # envs per H2D request
    host_envs = min(B, max(chunk, ((2 << 30) // per_env) // chunk * chunk))  # pinned pool: <= 2 GiB per rank, cycled
    free_host = psutil.virtual_memory().available
    free_dev, _ = th.cuda.mem_get_info(dev)
    if host_envs * per_env * world * 2 > free_host or bytes_in * 1.3 > free_dev:
        return {"value": None, "unit": "agent-steps/s", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": 0,
                "skipped": f"not enough memory (host free {free_host >> 30} GiB, device free {free_dev >> 30} GiB)"}
    lib = _lib.load()
    host = th.empty(host_envs, n, m, T, dtype=th.float32, pin_memory=True)
    for b0 in range(0, host_envs, chunk):  # fill the pinned pool with distinct synthetic values
        b1 = min(host_envs, b0 + chunk)
        host[b0:b1].copy_(th.rand(b1 - b0, n, m, T, device=dev))
    planes2 = th.empty_like(planes)
    staging = [th.empty(chunk * n * m * T, dtype=th.float32, device=dev) for _ in range(2)]
    copy_stream = th.cuda.Stream(device=dev)
    ret_host, act_host, rew_host, bytes_out = _d2h_buffers(th, runner, B, T, n)
    bufs = [planes, planes2]
    ready = [th.cuda.Event(), th.cuda.Event()]
    consumed = [th.cuda.Event(), th.cuda.Event()]

    def upload(dst, slot):
        with th.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])  # the rollout that read this buffer has finished
            for i, b0 in enumerate(range(0, B, chunk)):
                b1 = min(B, b0 + chunk)
                h0 = b0 % host_envs  # the pinned pool is cycled when it is smaller than the batch
                _lib.check(lib.sap_benefit_upload_host(host[h0:h0 + (b1 - b0)].data_ptr(), staging[i % 2].data_ptr(), dst[b0:b1].data_ptr(),
                                                       b1 - b0, n, m, T, copy_stream.cuda_stream), "sap_benefit_upload_host")
            ready[slot].record(copy_stream)

    def e2e_step(i):
        slot = i % 2
        th.cuda.current_stream().wait_event(ready[slot])
        runner.env.set_planes(bufs[slot])
        upload(bufs[1 - slot], 1 - slot)  # next episode's benefits stream in while this one rolls out
        with th.no_grad():
            batch = runner.run(test_mode=False)
        consumed[slot].record()
        buffer.insert_episode_batch(batch)
        ret_host.copy_(runner.last_episode_returns, non_blocking=True)
        act_host.copy_(batch["actions"], non_blocking=True)
        rew_host.copy_(batch["rewards"], non_blocking=True)

    graph = runner.args.use_cuda_graph
    runner.args.use_cuda_graph = False  # the planes pointer alternates: captured graphs would need one capture per buffer
    consumed[0].record()
    consumed[1].record()
    upload(bufs[0], 0)
    steps = max(2, min(opts.steps, opts.e2e_steps))
    e2e_step(0)  # warm-up episode (also primes the pipeline)
    th.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(1, steps + 1):
        e2e_step(i)
    e1.record()
    th.cuda.synchronize()
    ms = th.tensor([e0.elapsed_time(e1)], dtype=th.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    runner.env.set_planes(planes)
    runner.args.use_cuda_graph = graph
    val = world * B * n * T * steps / (ms.item() * 1e-3)
    return {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": bytes_out,
            "steps": steps, "ms_per_step": ms.item() / steps,
            "note": "benefit upload of episode e+1 overlaps the rollout of episode e (copy stream); every episode uploads "
                    f"all {B} envs' benefits from a pinned pool of {host_envs} envs ({host_envs * per_env >> 20} MiB, cycled); "
                    "eager launches (the planes buffer alternates)"}


# ----------------------------------------------------------------------------------------------- CPU baselines
def reference_throughput(w, weights, target_seconds, procs=None):
    """agent-steps/s of the UNMODIFIED reference (oracle/_ref or /root/reference): ParallelRunner with one env process per
    host core + BasicMAC + RNNAgent + epsilon-greedy, th.set_num_threads(1).  Bounded sample: the
    same n, m, L, M, N with a shorter horizon T_s (per-timestep work does not depend on T)."""
    from oracle import ref_runner as RR

    procs = procs or RR.host_cores()
    T_s = min(w["T"], 12 if w["n"] * w["m"] >= 2500 else 40)
    name, env_args = RR.reference_env_args(w, T=T_s, seed=7)
    t0 = time.perf_counter()
    v, sec, per_ep = RR.time_reference_rollout(name, env_args, procs, weights=weights, episodes=1, warmup=1)
    episodes = int(max(1, min(5, (target_seconds - (time.perf_counter() - t0)) / max(sec, 1e-3))))
    if episodes > 1:
        v, sec, per_ep = RR.time_reference_rollout(name, env_args, procs, weights=weights, episodes=episodes, warmup=1)
    sample = (f"reference ParallelRunner, {procs} env processes x T={T_s} timesteps of {w['n']}x{w['m']} ({w['env']} env), "
              f"BasicMAC + rnn agent(hidden 64) + epsilon_greedy, torch threads 1, median of {episodes} runner.run() calls")
    return v, procs, sample, time.perf_counter() - t0


def cpu_baseline(w, weights, seconds):
    from oracle import ref_runner as RR

    out = None
    if RR.reference_available():
        try:
            v, cores, sample, _ = reference_throughput(w, weights, seconds)
            out = {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": "reference", "sample": sample}
            # also the reference's EpisodeRunner on one core and the reference env alone
            # (step + get_pretransition_data, no agent, no buffer) for the kernel-vs-kernel comparison
            T_s = min(w["T"], 12 if w["n"] * w["m"] >= 2500 else 40)
            name, env_args = RR.reference_env_args(w, T=T_s, seed=7)
            try:
                out["episode_runner_1core"] = {"value": RR.time_reference_episode_runner(name, env_args, weights=weights),
                                               "unit": "agent-steps/s", "cores": 1,
                                               "sample": f"reference EpisodeRunner, one process, T={T_s}, median of 2 episodes"}
                es = RR.time_reference_env_only(name, env_args)
                out["env_only_1core"] = {"value": es, "unit": "env-steps/s", "agent_steps_per_s": es * w["n"], "cores": 1,
                                         "sample": f"reference env.step + get_pretransition_data, random actions, T={T_s}, 2 episodes"}
            except Exception as e:
                out["episode_runner_1core"] = {"value": None, "error": f"{type(e).__name__}: {e}"[:200]}
        except Exception as e:
            out = None
            err = f"{type(e).__name__}: {e}"[:200]
    pv, pcores, psample, _ = cpu_port_throughput(w, weights, target_seconds=min(seconds, 8.0))
    port = {"value": pv, "unit": "agent-steps/s", "cores": pcores, "kind": "port", "sample": psample}
    if out is None:
        out = dict(port)
        out["reference_unavailable"] = "no reference tree (oracle/_ref missing: run python oracle/make_ref.py)" if not RR.reference_available() else err
    else:
        out["port"] = port
    return out


# ----------------------------------------------------------------------------------------------- reference arm
def reference_arm(opts, w):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import ref_runner as RR

    weights = default_agent_weights(w)
    vals = []
    n_calls = opts.warmup + opts.steps
    kind = "reference" if RR.reference_available() else "port"
    for i in range(n_calls):
        per_step = min(opts.cpu_seconds, 150.0 / max(1, n_calls))  # whole run within a few minutes
        if kind == "reference":
            v, cores, sample, wall = reference_throughput(w, weights, per_step)
        else:
            v, cores, sample, wall = cpu_port_throughput(w, weights, target_seconds=per_step / 2)
        if i >= opts.warmup:
            vals.append((v, wall))
    vals.sort()
    v = vals[len(vals) // 2][0]
    T = w["T"]
    note = ("the UNMODIFIED reference (oracle/_ref, materialised by oracle/make_ref.py): ParallelRunner + env worker processes "
            "+ BasicMAC + RNNAgent + EpsilonGreedyActionSelector through its own API on the host cores"
            if kind == "reference" else
            "reference tree not available on this box: the numpy oracle port of the same path on all host cores")
    line = {"impl": "reference", "metric": "agent_steps_per_sec", "value": v, "unit": "agent-steps/s", "n_gpus": world,
            "steps": opts.steps, "warmup": opts.warmup, "ms_per_step": 1e3 * sum(x[1] for x in vals) / len(vals),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(opts, w, world), "note": note,
            "cpu_baseline": {"value": v, "unit": "agent-steps/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--envs-per-gpu", type=int, default=None)
    ap.add_argument("--selector", default="epsilon_greedy", choices=["epsilon_greedy", "sap"],
                    help="action selector of the rollout (sap = noise-perturbed optimal assignment per env; no CPU leg)")
    ap.add_argument("--graph", action="store_true", help="(default) replay the T-step loop of every episode as one CUDA graph")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel of every timestep eagerly")
    ap.add_argument("--agent-fc1", default="fp32", choices=["fp32", "fp16_split"],
                    help="first layer of the agent in the headline run (the other one is measured as a variant)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-fresh-benefits", action="store_true", help="also time the round-1 e2e definition (all benefits re-uploaded per episode)")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block of multi-GPU runs")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    opts = ap.parse_args()
    opts.warmup = max(opts.warmup, 3) if opts.impl == "ours" else max(opts.warmup, 0)
    w = dict(WORKLOADS[opts.workload])
    if opts.envs_per_gpu:
        w["B"] = opts.envs_per_gpu
        w["desc"] = f"{w['B']} envs x {w['n']} agents x {w['m']} tasks per GPU"
    if opts.impl == "reference":
        reference_arm(opts, w)
    else:
        gpu_arm(opts, w)


if __name__ == "__main__":
    main()
