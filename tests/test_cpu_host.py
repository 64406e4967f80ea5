"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares, the ctypes
structs match the C header, and the host-side mirrors of the reference interface behave like the reference."""
import ctypes
import os
import re
import subprocess
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
HEADER = os.path.join(ROOT, "include", "marl_sap_b200.h")


def _header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sap_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from marl_sap_b200 import _build, _lib

    _build.build()
    names = _header_functions()
    assert len(names) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), f"{name} declared in include/marl_sap_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding table and header disagree"
    lib = _lib.load()
    assert lib.sap_abi_version() == 1
    assert lib.sap_last_error() is not None


def test_ctypes_structs_match_c_header(tmp_path):
    """sizeof/offsetof as seen by a C compiler == the ctypes mirror in marl_sap_b200/_lib.py."""
    from marl_sap_b200 import _lib

    prog = tmp_path / "abi.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "marl_sap_b200.h"\nint main(void){\n'
                    'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(SapEnvDims), sizeof(SapField), sizeof(SapBatchView),'
                    ' offsetof(SapField, env_stride), offsetof(SapField, dtype), offsetof(SapBatchView, agent_in),'
                    ' offsetof(SapEnvDims, shared_planes), sizeof(SapSelectArgs), offsetof(SapSelectArgs, seed),'
                    ' offsetof(SapSelectArgs, eps));\nreturn 0;}\n')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.SapEnvDims), ctypes.sizeof(_lib.SapField), ctypes.sizeof(_lib.SapBatchView),
            _lib.SapField.env_stride.offset, _lib.SapField.dtype.offset, _lib.SapBatchView.agent_in.offset,
            _lib.SapEnvDims.shared_planes.offset, ctypes.sizeof(_lib.SapSelectArgs), _lib.SapSelectArgs.seed.offset,
            _lib.SapSelectArgs.eps.offset]
    assert got == want


def test_argument_validation_without_gpu():
    """Entry points validate their arguments before touching the device: null pointers / bad dims -> error codes."""
    from marl_sap_b200 import _lib

    lib = _lib.load()
    view = _lib.SapBatchView()
    ok_dims = _lib.SapEnvDims(2, 4, 6, 5, 3, 2, 2, 0)
    assert lib.sap_real_reset(ok_dims, None, None, None, None, None, None, view, None, None, None) == -1  # view.obs null
    dims = _lib.SapEnvDims(2, 4, 3, 5, 3, 2, 2, 0)  # m < n
    view.obs.ptr = 16
    view.obs.dtype = _lib.SAP_F16
    assert lib.sap_real_reset(dims, None, None, None, None, None, None, view, None, None, None) == -3
    assert b"m >= n" in lib.sap_last_error()
    dims = _lib.SapEnvDims(2, 4, 9, 5, 3, 3, 2, 0)  # odd M
    assert lib.sap_real_reset(dims, None, None, None, None, None, None, view, None, None, None) == -3
    assert lib.sap_select_epsilon_greedy(None, None, 1, 1, 1, 0.1, None, 0, None, None, None, None, None, None) == -1
    assert lib.sap_buffer_insert(None, None, 16, 4, 0, 0, 1, None) == -1
    assert lib.sap_benefit_ingest(None, None, 1, 1, 1, 1, None) == -1
    with pytest.raises(RuntimeError, match="failed"):
        _lib.check(-2, "example")


def test_product_path_requires_cuda():
    if th.cuda.is_available():
        pytest.skip("CPU-only check")
    from marl_sap_b200.action_selectors import REGISTRY as sel
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    with pytest.raises(RuntimeError, match="CUDA"):
        BatchedRealConstellationEnv(1, 4, 6, 3, 2, 2, 2, 0.5, sat_prox_mat=np.zeros((4, 6, 3), np.float32))
    s = sel["epsilon_greedy"](SimpleNamespace(epsilon_start=1.0, epsilon_finish=0.1, epsilon_anneal_time=10,
                                              evaluation_epsilon=0.0))
    with pytest.raises(RuntimeError, match="no CPU path"):
        s.select_action(th.zeros(1, 2, 3), th.ones(1, 2, 3, dtype=th.bool), 0)


def test_registries_have_the_reference_keys():
    from marl_sap_b200.action_selectors import REGISTRY as sel
    from marl_sap_b200.envs import REGISTRY as envs
    from marl_sap_b200.runners import REGISTRY as runners

    # /root/reference/src/action_selectors/__init__.py:9-18, envs/__init__.py:23-30, runners/__init__.py:1-7
    assert sorted(sel) == sorted(["continuous", "multinomial", "epsilon_greedy", "soft_policies", "sap",
                                  "epsilon_greedy_sap_test", "filtered_const_sap", "filtered_const_epsilon_greedy",
                                  "filtered_const_epsgr_sap_test", "filtered_const_soft_policies"])
    assert sorted(envs) == sorted(["benefit_obs_env", "mock_constellation_env", "power_constellation_env",
                                   "real_constellation_env", "real_power_constellation_env",
                                   "interference_constellation_env", "dictator_env"])
    assert sorted(runners) == ["episode", "parallel"]
    assert all(isinstance(c, type) and "Unavailable" not in c.__name__ for c in sel.values())  # all ten are built
    assert sel["sap"].__name__ == "SequentialAssignmentProblemSelector"
    with pytest.raises(NotImplementedError, match="hot path"):
        envs["dictator_env"]()
    with pytest.raises(NotImplementedError, match="sat_prox_mat"):
        envs["real_constellation_env"](num_planes=1, num_sats_per_plane=4, m=4, T=2, N=2, M=2, L=2, lambda_=0.5)


def test_epsilon_schedule_matches_reference_golden():
    from marl_sap_b200.components.epsilon_schedules import DecayThenFlatSchedule

    g = dict(np.load(os.path.join(GOLDEN, "selectors.npz")))
    s = DecayThenFlatSchedule(float(g["eps_start"]), float(g["eps_finish"]), float(g["eps_anneal"]), decay="linear")
    for t, e in zip(g["sched_t"], g["sched_eps"]):
        assert s.eval(int(t)) == e
    e = DecayThenFlatSchedule(1.0, 0.05, 100, decay="exp")
    assert e.eval(0) == 1.0 and 0.05 <= e.eval(10 ** 6) <= 0.0500001


def test_episode_batch_host_container_matches_reference_golden():
    """EpisodeBatch.update / ReplayBuffer ring semantics on CPU tensors (host logic) vs tests/golden/buffer.npz."""
    from marl_sap_b200.components.episode_buffer import EpisodeBatch, ReplayBuffer
    from marl_sap_b200.components.transforms import OneHot

    g = dict(np.load(os.path.join(GOLDEN, "buffer.npz")))
    n, m, L, T = (int(g[k]) for k in "nmLT")
    scheme = {
        "obs": {"vshape": 6, "group": "agents", "dtype": th.float16},
        "actions": {"vshape": (1,), "group": "agents", "dtype": th.int16},
        "avail_actions": {"vshape": (m,), "group": "agents", "dtype": th.bool},
        "rewards": {"vshape": (n,), "dtype": th.float16},
        "terminated": {"vshape": (1,), "dtype": th.bool},
        "prev_assigns": {"vshape": (n,), "dtype": th.int16, "part_of_state": True},
        "beta": {"vshape": (n, m, L), "dtype": th.float16, "part_of_state": True},
    }
    groups = {"agents": n}
    pre = {"actions": ("actions_onehot", [OneHot(out_dim=m)])}
    rb = ReplayBuffer(dict(scheme), groups, 5, T + 1, preprocess=pre, device="cpu")
    for e in range(4):
        B = 2
        batch = EpisodeBatch(dict(scheme), groups, B, T + 1, preprocess=pre, device="cpu")
        for t in range(T + 1):
            batch.update({"obs": [g[f"raw{e}_obs"][b, t] for b in range(B)],
                          "beta": [g[f"raw{e}_beta"][b, t] for b in range(B)],
                          "avail_actions": [[[1] * m] * n for b in range(B)],
                          "prev_assigns": [np.arange(n) for b in range(B)]}, ts=t)
            if t < T:
                batch.update({"actions": th.tensor(g[f"raw{e}_actions"][:, t]),
                              "rewards": [(list(g[f"raw{e}_rewards"][b, t]),) for b in range(B)],
                              "terminated": [(t == T - 1,) for b in range(B)]}, ts=t)
        for k, v in batch.data.transition_data.items():
            np.testing.assert_array_equal(v.numpy(), g[f"ep{e}_{k}"], err_msg=f"episode {e} field {k}")
        rb.insert_episode_batch(batch)
    for k, v in rb.data.transition_data.items():
        np.testing.assert_array_equal(v.numpy(), g[f"rb_{k}"], err_msg=f"replay field {k}")
    assert rb.buffer_index == int(g["rb_buffer_index"]) and rb.episodes_in_buffer == int(g["rb_episodes_in_buffer"])
    assert int(rb.max_t_filled()) == int(g["max_t_filled"])
    assert rb.can_sample(5) and not rb.can_sample(6)
    sub = rb[(("obs", "actions"))]
    assert set(sub.scheme) == {"obs", "actions"} and sub["obs"].shape == rb["obs"].shape
    sl = rb[1:3, 0:T]
    assert sl.batch_size == 2 and sl.max_seq_length == T
    with pytest.raises(KeyError):
        rb.update({"nope": [0]}, ts=0)
    with pytest.raises(ValueError):
        rb.update({"obs": [np.zeros((n, 7))]}, bs=0, ts=0)  # unsafe reshape
    with pytest.raises(IndexError):
        rb[:, [0, 1]]
    # lazy fields are rebuilt on access
    lazy = EpisodeBatch(dict(scheme), groups, 2, T + 1, preprocess=pre, device="cpu", lazy=("avail_actions", "actions_onehot"))
    assert "avail_actions" not in lazy.data.transition_data
    lazy.update({"actions": th.tensor([[1, 2, 0], [4, 4, 3]])}, ts=1)
    assert bool(lazy["avail_actions"].all()) and lazy["avail_actions"].shape == (2, T + 1, n, m)
    oh = lazy["actions_onehot"]
    assert oh.dtype == th.int16 and oh[0, 1].argmax(-1).tolist() == [1, 2, 0] and int(oh[:, 1].sum()) == 6


def test_cpu_oracle_port_throughput_helper_runs():
    """bench.py's CPU leg on a tiny workload (a few hundred ms)."""
    sys.path.insert(0, ROOT)
    import bench

    w = dict(bench.WORKLOADS["tiny"])
    weights = bench.default_agent_weights(w)
    dt, steps = bench._cpu_worker((w, 2, 3, 0, weights))
    assert steps == 2 * w["n"] * 3 and dt > 0
    assert bench.algorithmic_bytes_per_env_step(bench.WORKLOADS["c3"], 2) == 100 * 100 * 3 * 4 + 100 * 490 * 6 + 1616


def test_rollout_diagnostics_on_cpu_tensors():
    """utils/rollout_stats (device array ops replacing q_learner.py:157-191's loops) vs the loops, on CPU tensors."""
    import torch as th

    from marl_sap_b200.utils.rollout_stats import calc_conflicting_actions, calc_raw_benefits
    from oracle import cpu_oracle as O

    rng = np.random.default_rng(8)
    B, T, n, m = 3, 5, 7, 4
    acts = rng.integers(0, m, size=(B, T, n))
    beta = rng.random((B, T, n, m))
    a = th.tensor(acts).unsqueeze(-1)
    assert calc_conflicting_actions(a, m) == pytest.approx(O.calc_conflicting_actions(acts, m), rel=1e-12)
    assert calc_raw_benefits(th.tensor(beta), a) == pytest.approx(O.calc_raw_benefits(beta, acts), rel=1e-12)
    with pytest.raises(ValueError, match="unexpected shape"):
        calc_raw_benefits(th.tensor(beta[0]), a)
