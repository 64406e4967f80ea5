"""Drive the UNMODIFIED reference rollout (ParallelRunner / EpisodeRunner + BasicMAC + RNNAgent + epsilon-greedy) on the
host cores.

TEST / BASELINE INFRASTRUCTURE ONLY: imported by ``bench.py`` (``--impl reference`` and the ``cpu_baseline`` leg), by
``tests/golden/make_golden.py`` and by tests that are skipped when no reference tree is available.  The reference is
found by ``oracle/ref_import.py`` (``/root/reference`` in the build container, ``oracle/_ref`` on the GPU box - see
``oracle/make_ref.py``) and imported behind its stubs; none of this repo's kernels, envs, selectors or buffers are on
this path.

Follows BASELINE.md section 2 / SURVEY.md 8(d): ``batch_size_run`` = number of host cores, ``th.set_num_threads(1)``
like main.py:90, ``rnn`` agent (hidden 64, ``use_rnn=False``), ``epsilon_greedy``, the same n, m, T, L, N, M and a
``sat_prox_mat`` handed to every env; only ``runner.run()`` is timed (process start-up excluded).
"""
from __future__ import annotations

import contextlib
import io
import os
import time
from types import SimpleNamespace

import numpy as np


class _Log:
    def __init__(self):
        self.stats = {}

    def log_stat(self, k, v, t):
        self.stats.setdefault(k, []).append((t, float(v)))


def reference_available() -> bool:
    from oracle import ref_import

    return ref_import.reference_available()


def make_args(env_name, env_args, batch_size_run, runner="parallel", seed=0, **kw):
    """The slice of the reference's merged yaml config that the rollout path reads (config/default.yaml + env + alg)."""
    ea = dict(env_args)
    ea.setdefault("seed", seed)
    base = dict(env=env_name, env_args=ea, batch_size_run=batch_size_run, use_mps_action_selection=True, device="cpu",
                mac="basic_mac", render=False, test_nepisode=batch_size_run, runner_log_interval=10 ** 12, runner=runner,
                agent="rnn", hidden_dim=64, use_rnn=False, agent_output_type="q", action_selector="epsilon_greedy",
                epsilon_start=0.5, epsilon_finish=0.5, epsilon_anneal_time=1, evaluation_epsilon=0.0, obs_agent_id=False,
                obs_last_action=False, seed=seed)
    base.update(kw)
    return SimpleNamespace(**base)


def build_reference_runner(args, weights=None):
    """(runner, mac, buffer, logger) of the reference, wired like run.run_sequential (run.py:107-186)."""
    import torch as th

    from oracle import ref_import

    R = ref_import.ref_modules()
    log = _Log()
    runner_cls = R.parallel_runner.ParallelRunner if args.runner == "parallel" else R.episode_runner.EpisodeRunner
    with contextlib.redirect_stdout(io.StringIO()):
        runner = runner_cls(args, log)
        env = runner.get_env()
    args.n, args.m, args.T = env.n, env.m, env.T
    groups = {"agents": args.n}
    buffer = R.episode_buffer.ReplayBuffer(env.scheme, groups, max(2, args.batch_size_run), env.T + 1,
                                           preprocess=env.preprocess, device="cpu")
    th.manual_seed(args.seed)
    mac = R.basic_controller.BasicMAC(buffer.scheme, groups, args)
    if weights is not None:
        mac.agent.load_state_dict({k: th.tensor(v) for k, v in weights.items()})
        if hasattr(mac, "selector_agent"):  # only built when use_mps_action_selection is off (basic_controller.py:69-75)
            mac.update_action_selector_agent()
    with contextlib.redirect_stdout(io.StringIO()):
        runner.setup(scheme=env.scheme, groups=groups, preprocess=env.preprocess, mac=mac)
    return runner, mac, buffer, log


def time_reference_rollout(env_name, env_args, procs, weights=None, episodes=3, warmup=1):
    """Median agent-steps/s of ``runner.run()`` of the reference's ParallelRunner with ``procs`` env processes.
    Returns (agent_steps_per_s, seconds_per_episode, agent_steps_per_episode)."""
    import torch as th

    th.set_num_threads(1)  # main.py:90
    args = make_args(env_name, env_args, procs)
    runner, mac, buffer, _ = build_reference_runner(args, weights)
    per_ep = procs * args.n * args.T
    times = []
    try:
        with th.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            for i in range(warmup + episodes):
                t0 = time.perf_counter()
                batch = runner.run(test_mode=False)
                buffer.insert_episode_batch(batch)
                dt = time.perf_counter() - t0
                if i >= warmup:
                    times.append(dt)
    finally:
        with contextlib.redirect_stdout(io.StringIO()):
            try:
                runner.close_env()
            except Exception:
                pass
    times.sort()
    med = times[len(times) // 2]
    return per_ep / med, med, per_ep


def time_reference_episode_runner(env_name, env_args, weights=None, episodes=2, warmup=1):
    """the reference's EpisodeRunner, one process, one core.  Returns agent-steps/s."""
    import torch as th

    th.set_num_threads(1)
    args = make_args(env_name, env_args, 1, runner="episode")
    runner, mac, buffer, _ = build_reference_runner(args, weights)
    per_ep = args.n * args.T
    times = []
    with th.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        for i in range(warmup + episodes):
            t0 = time.perf_counter()
            batch = runner.run(test_mode=False)
            buffer.insert_episode_batch(batch)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    times.sort()
    return per_ep / times[len(times) // 2]


def time_reference_env_only(env_name, env_args, episodes=2):
    """the reference env alone on one core - ``step(actions)`` + ``get_pretransition_data()`` with uniform
    random actions, no agent, no buffer (the kernel-vs-kernel comparison).  Returns env-steps/s."""
    from oracle import ref_import

    R = ref_import.ref_modules()
    with contextlib.redirect_stdout(io.StringIO()):
        env = R.envs.REGISTRY[env_name](**env_args)
    rng = np.random.default_rng(0)
    steps, total = 0, 0.0
    for _ in range(episodes):
        env.reset()
        done = False
        while not done:
            a = rng.integers(0, env.m, size=env.n)
            t0 = time.perf_counter()
            _, done, _ = env.step(a)
            env.get_pretransition_data()
            total += time.perf_counter() - t0
            steps += 1
    return steps / total


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0)) or 1
    except AttributeError:
        return os.cpu_count() or 1


def reference_env_args(w, T=None, seed=0):
    """env_args of the reference for a bench workload dict (bench.WORKLOADS): same n, m, L, N, M; U(0,1) benefits."""
    rng = np.random.default_rng(seed)
    T = int(T or w["T"])
    S = rng.random((w["n"], w["m"], T), dtype=np.float32).astype(np.float64)
    if w["env"] == "real":
        return "real_constellation_env", dict(num_planes=1, num_sats_per_plane=w["n"], m=w["m"], T=T, N=w["N"], M=w["M"],
                                              L=w["L"], lambda_=0.5, sat_prox_mat=S, graphs=1)
    return "mock_constellation_env", dict(n=w["n"], m=w["m"], T=T, L=w["L"], lambda_=0.5, sat_prox_mat=S)
