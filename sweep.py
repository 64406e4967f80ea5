#!/usr/bin/env python
"""Synthetic step-throughput sweep (BASELINE.json configs[4], SURVEY.md 8(d) shape C5):
envs B in {256 .. 65536} x agents = tasks in {10, 50, 100, 200, 500}, kernel-only.

For every grid point it times, with CUDA events on the launching stream and inputs resident in HBM,
  * the env step + observation kernel (real env: M = N = 10, L = 3, fp16 scheme + fp32 agent input; 10 x 10 uses the
    mock env, like the reference's own 10 x 10 configuration: the real env needs m >= M + M/2),
  * the masked eps-greedy selector kernel on q[B, n, m],
  * the batched linear-sum-assignment kernel of the "sap" selectors on q + z * std (at most 16 384 envs),
and reports env-steps/s, agent-steps/s and the algorithmic GB/s against the measured HBM peak.
With --cpu it also times the numpy oracle port of the same env step on one host core for a few envs per shape (this is
the one place besides bench.py / tests where oracle/ is executed: as the CPU baseline, never as the product).

    python sweep.py [--out profiles/r01_sweep_c5.jsonl] [--cpu] [--quick]

B is capped so that one grid point's working set stays under --mem-gb (the cap is reported as `B_run`).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def real_obs(M, N, L):
    return M * L + N * M * L + N * (M // 2) * L + M


def time_launches(th, fn, reps):
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    th.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_point(th, lib_mod, B, n, opts, peak):
    lib = lib_mod.load()
    m, L, Tp = n, 3, 6
    mock = n < 16
    M = N = 0 if mock else 10
    obs = (L + 1) * m if mock else real_obs(M, N, L)
    e_obs = 4 if mock else 2
    dims1 = lib_mod.SapEnvDims(1, n, m, Tp, L, M, N, 0)
    scratch1 = 0 if mock else int(lib.sap_real_scratch_doubles(C.byref(dims1))) * 8
    per_env = Tp * n * m * 4 + n * obs * (e_obs + 4) + scratch1 + n * m * 4 + n * 64
    B_run = max(1, min(B, int(opts.mem_gb * 1e9 // per_env)))
    dev = "cuda"
    g = th.Generator(device=dev).manual_seed(n * 1000 + 1)
    planes = th.rand(B_run, Tp, n, m, device=dev, generator=g)
    stats = th.empty(B_run, Tp, 2, device=dev)
    lib_mod.check(lib.sap_benefit_stats(planes.data_ptr(), stats.data_ptr(), B_run, n, m, Tp, lib_mod.stream_ptr()), "stats")
    k = th.zeros(B_run, dtype=th.int32, device=dev)
    prev = th.zeros(B_run, n, dtype=th.int32, device=dev)
    ret = th.zeros(B_run, dtype=th.float64, device=dev)
    counts = th.zeros(B_run, m, dtype=th.int32, device=dev)
    top = th.zeros(B_run, n, max(M, 1), dtype=th.int32, device=dev)
    odt = th.float32 if mock else th.float16
    idt = th.int64 if mock else th.int16
    # one time slot aliased for every t (stride 0): the sweep times the kernel, it does not keep the episode
    obs_t = th.empty(B_run, 1, n, obs, dtype=odt, device=dev).expand(B_run, Tp + 1, n, obs)
    rew_t = th.empty(B_run, 1, n, dtype=odt, device=dev).expand(B_run, Tp + 1, n)
    act_t = th.empty(B_run, 1, n, 1, dtype=idt, device=dev).expand(B_run, Tp + 1, n, 1)
    term_t = th.empty(B_run, 1, 1, dtype=th.uint8, device=dev).expand(B_run, Tp + 1, 1)
    fill_t = th.empty(B_run, 1, 1, dtype=th.int64, device=dev).expand(B_run, Tp + 1, 1)
    pa_t = th.empty(B_run, 1, n, dtype=idt, device=dev).expand(B_run, Tp + 1, n)
    ain_mode = "f32" if mock else opts.agent_in
    ain = th.empty(B_run, n, obs, dtype=th.float16 if ain_mode == "f16" else th.float32, device=dev)
    view = lib_mod.SapBatchView()
    view.obs, view.rewards, view.actions = lib_mod.field_of(obs_t), lib_mod.field_of(rew_t), lib_mod.field_of(act_t)
    view.terminated, view.filled = lib_mod.field_of(term_t), lib_mod.field_of(fill_t)
    if not mock:
        view.prev_assigns = lib_mod.field_of(pa_t)
    f = lib_mod.SapField()
    f.ptr, f.env_stride, f.t_stride = ain.data_ptr(), n * obs, obs
    f.dtype = lib_mod.SAP_F16 if ain_mode == "f16" else lib_mod.SAP_F32
    if ain_mode != "none":
        view.agent_in = f
    dims = lib_mod.SapEnvDims(B_run, n, m, Tp, L, M, N, 0)
    scratch = None
    if not mock:
        need = int(lib.sap_real_scratch_doubles(C.byref(dims)))
        scratch = th.empty(need, dtype=th.float64, device=dev) if need else None
    acts = th.randint(0, m, (B_run, n), device=dev, generator=g)
    st = lib_mod.stream_ptr()
    if mock:
        prev0 = th.stack([th.randperm(m, device=dev)[:n] for _ in range(min(B_run, 64))])
        prev0 = prev0.repeat((B_run + prev0.shape[0] - 1) // prev0.shape[0], 1)[:B_run].contiguous()

        def reset():
            lib_mod.check(lib.sap_mock_reset(C.byref(dims), planes.data_ptr(), prev0.data_ptr(), k.data_ptr(), prev.data_ptr(),
                                             ret.data_ptr(), C.byref(view), st), "mock_reset")

        def step():
            lib_mod.check(lib.sap_mock_step(C.byref(dims), planes.data_ptr(), None, 0.5, acts.data_ptr(), k.data_ptr(),
                                            prev.data_ptr(), ret.data_ptr(), counts.data_ptr(), C.byref(view), st), "mock_step")
    else:
        def reset():
            lib_mod.check(lib.sap_real_reset(C.byref(dims), planes.data_ptr(), stats.data_ptr(), None, k.data_ptr(),
                                             prev.data_ptr(), ret.data_ptr(), C.byref(view), top.data_ptr(),
                                             lib_mod.ptr(scratch), st), "real_reset")

        def step():
            lib_mod.check(lib.sap_real_step(C.byref(dims), planes.data_ptr(), stats.data_ptr(), None, None, 0.5,
                                            acts.data_ptr(), k.data_ptr(), prev.data_ptr(), ret.data_ptr(), counts.data_ptr(),
                                            C.byref(view), top.data_ptr(), lib_mod.ptr(scratch), st), "real_step")

    full = Tp - L  # steps whose new window still has L planes
    ms = []
    for rnd in range(opts.rounds + 1):  # round 0 = warm-up
        reset()
        t = time_launches(th, step, full)
        if rnd:
            ms.append(t)
    env_ms = sorted(ms)[len(ms) // 2]
    e_ain = {"f32": 4, "f16": 2, "none": 0}[ain_mode]
    bytes_step = n * m * L * 4 + n * obs * (e_obs + e_ain) + n * 16 + 16
    bytes_8d = n * m * L * 4 + n * obs * 4 + n * 16 + 16   # window + fp32 obs + scalars
    rec = {"B": B, "B_run": B_run, "n": n, "m": m, "env": "mock" if mock else "real", "obs_size": obs,
           "env_kernel_ms": round(env_ms, 5), "env_steps_per_s": B_run / env_ms * 1e3,
           "agent_steps_per_s": B_run * n / env_ms * 1e3,
           "algorithmic_bytes_per_env_step": bytes_step,
           "hbm_gbps": B_run * bytes_step / env_ms / 1e6, "hbm_frac": B_run * bytes_step / env_ms / 1e6 / peak,
           "agent_in": ain_mode, "hbm_frac_8d": B_run * bytes_8d / env_ms / 1e6 / peak}
    del planes, obs_t, ain, scratch
    th.cuda.empty_cache()
    if opts.env_only:
        return rec
    # selector: masked eps-greedy over q[B, n, m] (all actions available), Philox draws in-kernel
    Bs = max(1, min(B_run, int(opts.mem_gb * 1e9 // (n * m * 5 + n * 8))))
    q = th.randn(Bs, n, m, device=dev, generator=g)
    out = th.empty(Bs, n, dtype=th.int64, device=dev)
    ks = th.zeros(Bs, dtype=th.int32, device=dev)

    def select():
        lib_mod.check(lib.sap_select_epsilon_greedy(q.data_ptr(), None, Bs, n, m, 0.05, None, 1234, None, ks.data_ptr(), None,
                                                    None, out.data_ptr(), st), "select")

    select()
    for _ in range(3):
        select()
    sel_ms = time_launches(th, select, 20)
    sel_bytes = n * m * 4 + n * 8
    rec.update({"select_B_run": Bs, "select_kernel_ms": round(sel_ms, 5),
                "select_hbm_gbps": Bs * sel_bytes / sel_ms / 1e6, "select_hbm_frac": Bs * sel_bytes / sel_ms / 1e6 / peak})
    # assignment selector: one linear-sum-assignment per env on q + z * std (the "sap" selector's kernel)
    Bl = max(1, min(Bs, int(opts.mem_gb * 1e9 // (n * m * 8 + n * 8)), 16384))
    z = th.randn(Bl, n, m, device=dev, generator=g)
    std = (q[:Bl].abs().mean(dim=(1, 2)) * 0.1).contiguous()
    cols = th.empty(Bl, n, dtype=th.int64, device=dev)

    def lsa():
        lib_mod.check(lib.sap_lsa_maximize(q.data_ptr(), z.data_ptr(), std.data_ptr(), Bl, n, m, cols.data_ptr(), None, st),
                      "lsa")

    lsa()
    lsa_ms = time_launches(th, lsa, 3)
    rec.update({"lsa_B_run": Bl, "lsa_kernel_ms": round(lsa_ms, 5), "lsa_us_per_env": lsa_ms / Bl * 1e3})
    del q, out, z, cols
    th.cuda.empty_cache()
    return rec


def cpu_point(n, seconds=4.0):
    """numpy oracle port of the same env step on ONE host core (bounded sample: a few envs, a few steps)."""
    import numpy as np
    from oracle import cpu_oracle as O

    m, L, T = n, 3, 6
    rng = np.random.default_rng(n)
    Bc = 4 if n <= 100 else 1
    S = rng.random((Bc, n, m, T), dtype=np.float32).astype(np.float64)
    if n < 16:
        st = O.MockState(S, L, 0.5)
        st.reset(np.stack([rng.permutation(m)[:n] for _ in range(Bc)]))
    else:
        st = O.RealState(S, L, 10, 10, 0.5)
        st.reset()
    steps, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds and steps < T - L:
        st.step(rng.integers(0, m, size=(Bc, n)))  # includes the observation build, like the kernel
        st.pretransition()
        steps += 1
    dt = time.perf_counter() - t0
    return {"cpu_envs": Bc, "cpu_steps": steps, "cpu_cores": 1, "cpu_env_steps_per_s": Bc * steps / dt,
            "cpu_agent_steps_per_s": Bc * steps * n / dt}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--ns", type=str, default=None, help="comma list of n = m values (default: the C5 grid)")
    ap.add_argument("--Bs", type=str, default=None, help="comma list of env counts (default: the C5 grid)")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--env-only", action="store_true", help="time the env kernel only (kernel tuning runs)")
    ap.add_argument("--kernel-path", type=int, default=0,
                    help="real-env kernel override (sap_real_select_kernel): 0 auto, 1 generic, 2 / 3 multi-CTA, 4 first-generation")
    ap.add_argument("--agent-in", default="f32", choices=["f32", "f16", "none"],
                    help="agent-input staging the env kernel writes next to the obs rows")
    ap.add_argument("--mem-gb", type=float, default=80.0)
    opts = ap.parse_args()
    import torch as th

    from marl_sap_b200 import _lib as lib_mod

    if not th.cuda.is_available():
        raise SystemExit("sweep.py needs a CUDA device (there is no CPU path)")
    if opts.kernel_path:
        assert lib_mod.load().sap_real_select_kernel(opts.kernel_path) >= 0
    peak, peak_src = hbm_peak()
    Bs = [256, 4096] if opts.quick else [256, 1024, 4096, 16384, 65536]
    ns = [10, 100] if opts.quick else [10, 50, 100, 200, 500]
    if opts.ns:
        ns = [int(x) for x in opts.ns.split(",")]
    if opts.Bs:
        Bs = [int(x) for x in opts.Bs.split(",")]
    recs = []
    cpu = {n: cpu_point(n) for n in ns} if opts.cpu else {}
    for n in ns:
        for B in Bs:
            rec = run_point(th, lib_mod, B, n, opts, peak)
            rec.update(cpu.get(n, {}))
            rec["peak_gbps"], rec["peak_source"] = peak, peak_src
            recs.append(rec)
            print(json.dumps(rec), flush=True)
    if opts.out:
        with open(opts.out, "w") as f:
            for r in recs:
                f.write(json.dumps(r) + "\n")
    print("\n| env | n = m | B (run) | env kernel ms | env-steps/s | agent-steps/s | GB/s | of HBM peak | select ms | select of peak | LSA us/env |"
          + (" 1-core numpy env-steps/s |" if opts.cpu else ""))
    print("|---|---|---|---|---|---|---|---|---|---|---|" + ("---|" if opts.cpu else ""))
    for r in recs:
        if opts.env_only:
            print(f"| {r['env']} | {r['n']} | {r['B']} ({r['B_run']}) | {r['env_kernel_ms']:.4f} | {r['env_steps_per_s']:.3g} | "
                  f"{r['agent_steps_per_s']:.3g} | {r['hbm_gbps']:.0f} | {r['hbm_frac']:.2f} |")
            continue
        row = (f"| {r['env']} | {r['n']} | {r['B']} ({r['B_run']}) | {r['env_kernel_ms']:.4f} | {r['env_steps_per_s']:.3g} | "
               f"{r['agent_steps_per_s']:.3g} | {r['hbm_gbps']:.0f} | {r['hbm_frac']:.2f} | {r['select_kernel_ms']:.4f} | "
               f"{r['select_hbm_frac']:.2f} | {r['lsa_us_per_env']:.2f} |")
        if opts.cpu:
            row += f" {r['cpu_env_steps_per_s']:.3g} |"
        print(row)


if __name__ == "__main__":
    main()
