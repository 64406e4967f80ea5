"""GPU parity of the env kernels (through the C ABI) against the golden vectors recorded from the
reference and against the CPU oracle on seeded inputs.  Bit-exact for indices/flags; obs/beta values are
gathers (exact after the single rounding to the scheme dtype); rewards rounded once from float64."""
import copy
import glob
import os

import numpy as np
import pytest
import torch as th

from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REAL = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "real_*.npz"))) + ["kat1_real.npz"]
MOCK = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "mock_*.npz")))


def _load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def _batch_for(env, B, real_dtype=None, lazy=()):
    from marl_sap_b200.components.episode_buffer import EpisodeBatch

    scheme = copy.deepcopy(env.scheme)
    if real_dtype is not None:
        for k in ("obs", "rewards", "beta"):
            scheme[k]["dtype"] = real_dtype
    return EpisodeBatch(scheme, {"agents": env.n}, B, env.T + 1, preprocess=env.preprocess, device="cuda", lazy=lazy)


def _cast(x64, dtype):
    """The reference's single rounding th.tensor(np.array(v), dtype=...) (episode_buffer.py:107-108)."""
    return th.tensor(np.asarray(x64), dtype=dtype)


@pytest.mark.parametrize("name", REAL)
@pytest.mark.parametrize("real_dtype", [th.float16, th.float32])
@pytest.mark.parametrize("shared", [True, False])
@pytest.mark.parametrize("generic", ["0", "1", "2", "3"])
def test_real_env_matches_reference_golden(name, real_dtype, shared, generic, real_kernel_path):
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    # "0": automatic (shared-memory fast path), "1": generic one-CTA-per-env kernel, "2" / "3": multi-CTA path of the large
    # shapes (keyed lists with certificates / exact float64 selection)
    real_kernel_path(generic)

    g = _load(name)
    B = 3
    S = g["S"]
    n, m, T = S.shape
    Sin = S if shared else np.broadcast_to(S, (B, n, m, T)).copy()
    env = BatchedRealConstellationEnv(B, n, m, T, int(g["L_arg"]), int(g["M"]), int(g["N"]), float(g["lambda_"]),
                                      sat_prox_mat=Sin, task_prios=g.get("task_prios"), T_ctor=int(g["T_ctor"]))
    assert env.L == int(g["L"]) and env.obs_size == int(g["obs_size"])
    batch = _batch_for(env, B, real_dtype)
    env.reset(batch)
    acts = g["actions"]
    for t in range(acts.shape[0]):
        a = th.tensor(np.broadcast_to(acts[t], (B, n)).copy(), device="cuda")
        done = env.step(a, batch)
        assert done == bool(g["done"][t])
    th.cuda.synchronize()
    td = {k: v.cpu() for k, v in batch.data.transition_data.items()}
    Tn = acts.shape[0]
    for b in range(B):
        assert th.equal(td["obs"][b], _cast(g["obs"], real_dtype)), f"obs mismatch env {b}"
        assert th.equal(td["beta"][b], _cast(g["beta"], real_dtype))
        assert th.equal(td["prev_assigns"][b], _cast(g["prev"], th.int16))
        assert th.equal(td["rewards"][b, :Tn], _cast(g["rewards"], real_dtype))
        assert th.equal(td["actions"][b, :Tn, :, 0], _cast(acts, th.int16))
        assert th.equal(td["actions_onehot"][b, :Tn], _cast(O.one_hot(acts, m, np.int16), th.int16))
        assert td["terminated"][b, :Tn, 0].tolist() == [bool(d) for d in g["done"]]
        assert td["filled"][b, :, 0].tolist() == [1] * (Tn + 1)
        assert bool(td["avail_actions"][b].all())
    # slots never written stay zero
    assert not td["rewards"][:, Tn:].any() and not td["actions"][:, Tn:].any() and not td["terminated"][:, Tn:].any()


@pytest.mark.parametrize("cfg", [
    dict(B=5, n=50, m=50, T=6, L=3, M=10, N=10, gen="dense", seed=0),
    dict(B=3, n=20, m=64, T=5, L=3, M=10, N=10, gen="ties", seed=1),
    dict(B=2, n=100, m=100, T=4, L=3, M=10, N=10, gen="dense", seed=2),
    dict(B=2, n=33, m=47, T=5, L=2, M=6, N=5, gen="ref", seed=3),
    dict(B=2, n=40, m=130, T=3, L=4, M=8, N=3, gen="ref", seed=4),
    dict(B=3, n=64, m=64, T=4, L=3, M=10, N=10, gen="dup", seed=5),
    dict(B=2, n=30, m=45, T=4, L=3, M=10, N=10, gen="neg", seed=6),
    dict(B=2, n=16, m=16, T=3, L=1, M=10, N=10, gen="dense", seed=7),
    dict(B=2, n=12, m=200, T=3, L=3, M=10, N=10, gen="const", seed=8),
])
@pytest.mark.parametrize("generic", ["0", "1", "2", "3"])
def test_real_env_matches_oracle(cfg, generic, real_kernel_path):
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    real_kernel_path(generic)
    rng = np.random.default_rng(cfg["seed"])
    B, n, m, T, L, M, N = (cfg[k] for k in ("B", "n", "m", "T", "L", "M", "N"))
    if cfg["gen"] == "dense":
        S = O.gen_dense(rng, B, n, m, T)
    elif cfg["gen"] == "ties":
        S = (np.round(O.gen_exact(rng, B, n, m, T, zero_frac=0.5) * 4) / 4).astype(np.float32)
    elif cfg["gen"] == "dup":      # duplicates above the minimum + near-ties one ulp apart: exercises the exact redo path
        S = (np.round(O.gen_dense(rng, B, n, m, T) * 16) / 16 + 1).astype(np.float32)
        S[:, ::3] = np.nextafter(S[:, ::3], np.float32(4))
    elif cfg["gen"] == "neg":      # negative benefits: lo < 0, zeros are not the minimum
        S = (O.gen_ref_like(rng, B, n, m, T) - np.float32(0.25) * (rng.random((B, n, m, T)) < 0.1)).astype(np.float32)
    elif cfg["gen"] == "const":    # every window sum identical
        S = np.full((B, n, m, T), 0.5, dtype=np.float32)
    else:
        S = O.gen_ref_like(rng, B, n, m, T)
    prios = (rng.integers(1, 4, size=m) * 0.5).astype(np.float32) if cfg["seed"] % 2 else None
    acts = rng.integers(0, m, size=(T, B, n))
    acts[:, :, : n // 3] = acts[:, :, :1]
    st = O.RealState(S.astype(np.float64), L, M, N, 0.5, task_prios=prios)
    want = O.rollout(st, lambda t, pre: acts[t], "real")
    env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5, sat_prox_mat=S, task_prios=prios)
    batch = _batch_for(env, B, th.float32)
    env.reset(batch)
    tops = [env.top.cpu().numpy().copy()]
    counts = []
    for t in range(T):
        env.step(th.tensor(acts[t], device="cuda"), batch)
        tops.append(env.top.cpu().numpy().copy())
        counts.append(env.counts.cpu().numpy().copy())
    td = {k: v.cpu() for k, v in batch.data.transition_data.items()}
    assert th.equal(td["obs"], _cast(want["obs"], th.float32))
    assert th.equal(td["beta"], _cast(want["beta"], th.float32))
    assert th.equal(td["rewards"], _cast(want["rewards"], th.float32))
    assert th.equal(td["prev_assigns"], _cast(want["prev_assigns"], th.int16))
    assert th.equal(td["terminated"][..., 0], th.tensor(want["terminated"]))
    assert th.equal(td["filled"][..., 0], th.tensor(want["filled"]))
    np.testing.assert_array_equal(np.stack(counts, 1), want["counts"][:, :T])
    # the env's shared top-M equals the oracle's stable top-M of every pre-transition beta
    for t in range(T):
        np.testing.assert_array_equal(tops[t], O.top_m_tasks(want["beta"][:, t], M))
    np.testing.assert_allclose(env.ep_return.cpu().numpy(), want["rewards"].sum((1, 2)), rtol=1e-12)


@pytest.mark.parametrize("name", MOCK)
def test_mock_env_matches_reference_golden(name):
    from marl_sap_b200.envs.batched import BatchedMockConstellationEnv

    g = _load(name)
    S = g["S"]
    n, m, T = S.shape
    B = 2
    env = BatchedMockConstellationEnv(B, n, m, T, int(g["L"]), float(g["lambda_"]), sat_prox_mat=S)
    batch = _batch_for(env, B)
    env.reset(batch, prev0=np.broadcast_to(g["prev0"], (B, n)).copy())
    acts = g["actions"]
    for t in range(T):
        done = env.step(th.tensor(np.broadcast_to(acts[t], (B, n)).copy(), device="cuda"), batch)
        assert done == bool(g["done"][t])
    td = {k: v.cpu() for k, v in batch.data.transition_data.items()}
    for b in range(B):
        assert th.equal(td["obs"][b], _cast(g["obs"], th.float32))
        assert th.equal(td["beta"][b], _cast(g["beta"], th.float32))
        assert th.equal(td["rewards"][b, :T], _cast(g["rewards"], th.float32))
        assert th.equal(td["actions"][b, :T, :, 0], th.tensor(acts))
        assert not td["prev_assigns"][b].any()  # the mock env never emits prev_assigns (mock_constellation_env.py:170-174)
        assert td["terminated"][b, :T, 0].tolist() == [bool(d) for d in g["done"]]
        assert td["filled"][b, :, 0].tolist() == [1] * (T + 1)


@pytest.mark.parametrize("cfg", [dict(B=4, n=10, m=10, T=7, L=3), dict(B=3, n=100, m=100, T=5, L=3),
                                 dict(B=2, n=17, m=23, T=4, L=2), dict(B=2, n=50, m=52, T=3, L=5),
                                 # tiny envs, many of them: the one-warp-per-env variant (8 envs per CTA)
                                 dict(B=70, n=10, m=10, T=7, L=3), dict(B=67, n=17, m=23, T=4, L=2),
                                 dict(B=130, n=10, m=12, T=5, L=3), dict(B=64, n=3, m=40, T=3, L=4)])
def test_mock_env_matches_oracle(cfg):
    from marl_sap_b200.envs.batched import BatchedMockConstellationEnv

    B, n, m, T, L = (cfg[k] for k in ("B", "n", "m", "T", "L"))
    rng = np.random.default_rng(B * 1000 + n)
    S = O.gen_ref_like(rng, B, n, m, T)
    prev0 = np.stack([rng.permutation(m)[:n] for _ in range(B)])
    acts = rng.integers(0, m, size=(T, B, n))
    acts[:, :, : n // 2] = acts[:, :, :1]
    Tt = rng.integers(0, 2, size=(m, m)).astype(np.float64) if n == 17 else None
    st = O.MockState(S.astype(np.float64), L, 0.5, T_trans=Tt)
    want = O.rollout(st, lambda t, pre: acts[t], "mock", prev0=prev0)
    env = BatchedMockConstellationEnv(B, n, m, T, L, 0.5, sat_prox_mat=S, T_trans=Tt)
    batch = _batch_for(env, B)
    env.reset(batch, prev0=prev0)
    for t in range(T):
        env.step(th.tensor(acts[t], device="cuda"), batch)
    td = {k: v.cpu() for k, v in batch.data.transition_data.items()}
    assert th.equal(td["obs"], _cast(want["obs"], th.float32))
    assert th.equal(td["beta"], _cast(want["beta"], th.float32))
    assert th.equal(td["rewards"], _cast(want["rewards"], th.float32))
    assert th.equal(td["terminated"][..., 0], th.tensor(want["terminated"]))
    assert th.equal(td["actions_onehot"][:, :T], th.tensor(O.one_hot(np.moveaxis(acts, 0, 1), m, np.int64)))


def test_real_env_custom_T_trans_and_lazy_fields():
    """T_trans override (real_constellation_env.py:66, :314) and the lazy buffer mode (beta rebuilt on access)."""
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    rng = np.random.default_rng(9)
    B, n, m, T, L, M, N = 2, 12, 20, 5, 3, 4, 3
    S = O.gen_ref_like(rng, B, n, m, T)
    Tt = rng.integers(0, 2, size=(m, m)).astype(np.float64)
    acts = rng.integers(0, m, size=(T, B, n))
    st = O.RealState(S.astype(np.float64), L, M, N, 0.5, T_trans=Tt)
    want = O.rollout(st, lambda t, pre: acts[t], "real")
    env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5, sat_prox_mat=S, T_trans=Tt)
    batch = _batch_for(env, B, lazy=("beta", "avail_actions", "actions_onehot"))
    batch.set_lazy_provider("beta", lambda _b: env.beta_field(th.float16))
    env.reset(batch)
    for t in range(T):
        env.step(th.tensor(acts[t], device="cuda"), batch)
    assert "beta" not in batch.data.transition_data and "avail_actions" not in batch.data.transition_data
    assert th.equal(batch["rewards"].cpu(), _cast(want["rewards"], th.float16))
    assert th.equal(batch["obs"].cpu(), _cast(want["obs"], th.float16))
    assert th.equal(batch["beta"].cpu(), _cast(want["beta"], th.float16))
    assert bool(batch["avail_actions"].all()) and batch["avail_actions"].shape == (B, T + 1, n, m)
    oh = batch["actions_onehot"].cpu()
    assert th.equal(oh[:, :T], _cast(O.one_hot(np.moveaxis(acts, 0, 1), m, np.int16), th.int16))


def test_real_env_large_shape_properties():
    """BASELINE-size envs (100 x 100, M = N = 10): size-independent properties instead of an oracle replay."""
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    B, n, m, T, L, M, N = 64, 100, 100, 8, 3, 10, 10
    g = th.Generator().manual_seed(0)
    S = th.rand(B, n, m, T, generator=g)
    env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5, sat_prox_mat=S.cuda())
    batch = _batch_for(env, B, th.float32, lazy=("beta", "avail_actions", "actions_onehot"))
    env.reset(batch)
    planes = env.planes.double()
    for t in range(T):
        acts = th.randint(0, m, (B, n), generator=g).cuda()
        env.step(acts, batch)
        # conflict counts: a histogram of the actions
        want_cnt = th.zeros(B, m, dtype=th.int32, device="cuda").scatter_add_(1, acts, th.ones_like(acts, dtype=th.int32))
        assert th.equal(env.counts, want_cnt)
        if t + 1 < T:
            tot = planes[:, t + 1:t + 1 + L].sum(1)  # [B,n,m] float64 window sums
            top = env.top.long()
            top_vals = tot.gather(2, top)
            assert bool((top_vals[..., :-1] >= top_vals[..., 1:]).all()), "top-M not sorted by window sum"
            rest = tot.scatter(2, top, float("-inf")).max(-1).values
            assert bool((rest <= top_vals[..., -1]).all()), "an excluded task beats the M-th task"
            obs = batch["obs"][:, t + 1]
            local = obs[..., : M * L].reshape(B, n, M, L)
            Leff = min(L, T - (t + 1))
            want_local = planes[:, t + 1:t + 1 + Leff].permute(0, 2, 3, 1).gather(
                2, top[..., None].expand(B, n, M, Leff)).float()
            assert th.equal(local[..., :Leff], want_local)
            flags = obs[..., -M:]
            assert th.equal(flags, (top == acts[..., None]).float())
    assert not batch["obs"][:, T].any()
    # determinism: a second identical rollout gives identical bytes
    first = batch["obs"].clone()
    g2 = th.Generator().manual_seed(0)
    _ = th.rand(B, n, m, T, generator=g2)
    env.reset(batch)
    for t in range(T):
        env.step(th.randint(0, m, (B, n), generator=g2).cuda(), batch)
    assert th.equal(first, batch["obs"])


def test_real_env_precondition_errors():
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    S = np.zeros((1, 4, 3, 2), dtype=np.float32)  # m < n
    env = BatchedRealConstellationEnv(1, 4, 3, 2, 1, 2, 2, 0.5, sat_prox_mat=S)
    batch = _batch_for(env, 1)
    with pytest.raises(RuntimeError, match="m >= n"):
        env.reset(batch)
    S = np.zeros((1, 4, 9, 2), dtype=np.float32)
    env = BatchedRealConstellationEnv(1, 4, 9, 2, 1, 3, 2, 0.5, sat_prox_mat=S)  # odd M
    with pytest.raises(RuntimeError, match="even"):
        env.reset(_batch_for(env, 1))
    with pytest.raises(ValueError):
        env.load_benefits(np.zeros((4, 9, 3), dtype=np.float32))


def test_single_env_facade_kat1():
    """REGISTRY['real_constellation_env'] object driven like the reference env (KAT-1, experiments.py:265-288)."""
    from marl_sap_b200.envs import REGISTRY

    g = _load("kat1_real.npz")
    env = REGISTRY["real_constellation_env"](num_planes=1, num_sats_per_plane=4, m=4, T=1, N=2, M=2, L=2, lambda_=0.5,
                                             sat_prox_mat=g["S"].astype(np.float64), graphs=1)
    assert (env.n, env.m, env.T, env.L) == (4, 4, 2, 1)
    env.reset()
    pre = env.get_pretransition_data()
    np.testing.assert_array_equal(np.array(pre["obs"][0]), g["obs"][0])
    np.testing.assert_array_equal(pre["prev_assigns"][0], np.arange(4))
    r, d, info = env.step([0, 0, 2, 3])
    np.testing.assert_allclose(r, [2.5, 0.75, 4.0, 10.0])
    assert d is False and info == {}
    import copy as _copy
    import pickle

    env2 = _copy.deepcopy(env)
    r, d, info = env.step([1, 1, 1, 3])
    np.testing.assert_allclose(r, [1 / 6, 1 / 6, 1 / 6, 1.0], rtol=1e-6)
    assert d is True and not np.array(env.get_obs()).any()
    r2, d2, _ = env2.step([1, 1, 1, 3])  # deep copy carries the live state (HAAL-style use)
    np.testing.assert_allclose(r2, r)
    env3 = pickle.loads(pickle.dumps(env))
    assert env3.n == 4 and env3._impl is None
    bh = env.beta_hat(g["beta"][0], np.arange(4))
    np.testing.assert_allclose(bh[..., 0], g["beta"][0][..., 0] - 0.5 * (1 - np.eye(4)) * (g["beta"][0].sum(-1) > 1e-12))


def test_real_env_full_size_fast_equals_generic(real_kernel_path):
    """BASELINE shape 100 x 100 (M = N = 10, L = 3, fp16 scheme) at a few hundred envs: both generations of the shared-memory
    kernel, the generic float64 kernel and the multi-CTA large-shape path (both modes) must produce identical bytes (obs,
    rewards, top-M, agent input) step after step."""
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    B, n, m, T, L, M, N = 296, 100, 100, 6, 3, 10, 10
    g = th.Generator().manual_seed(3)
    S = th.rand(B, n, m, T, generator=g)
    S[:, :, ::7] = 0.0            # inactive tasks: exact zero ties
    S[:40] = (S[:40] * 8).round() / 8  # coarse grid: duplicate sums above zero
    acts = [th.randint(0, m, (B, n), generator=g).cuda() for _ in range(T)]
    outs = []
    for generic in ("1", "2", "3", "4", "5", "0"):  # 5: the bench kernel reading n, m at run time; 0: its 100 x 100 instantiation
        real_kernel_path(generic)
        env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5, sat_prox_mat=S.cuda())
        batch = _batch_for(env, B, lazy=("beta", "avail_actions", "actions_onehot"))
        batch.agent_in = th.zeros(B, n, env.obs_size, device="cuda")
        env.reset(batch)
        ain = [batch.agent_in.clone()]
        tops = [env.top.clone()]
        for t in range(T):
            env.step(acts[t], batch)
            ain.append(batch.agent_in.clone())
            tops.append(env.top.clone())
        outs.append((batch["obs"].clone(), batch["rewards"].clone(), th.stack(ain), th.stack(tops[:-1]), env.ep_return.clone()))
    for fields in zip(*outs):
        assert all(th.equal(fields[0], other) for other in fields[1:])
    obs, _, ain, _, _ = outs[-1]
    assert th.equal(ain, obs.permute(1, 0, 2, 3).float())  # agent_in == float(obs[:, t]) for every t


@pytest.mark.parametrize("path", [0, 5])  # 0: instantiation with the shape compiled in; 5: shape read at run time
def test_real_env_constellation_scale_matches_oracle(path, real_kernel_path):
    """324 agents x 450 tasks (real_constellation_env.yaml): too large for one SM: runs the multi-CTA path."""
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    real_kernel_path(path)

    rng = np.random.default_rng(11)
    B, n, m, T, L, M, N = 2, 324, 450, 3, 3, 10, 10
    S = O.gen_ref_like(rng, B, n, m, T, p_active=0.05)
    acts = rng.integers(0, m, size=(T, B, n))
    st = O.RealState(S.astype(np.float64), L, M, N, 0.5)
    want = O.rollout(st, lambda t, pre: acts[t], "real")
    env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5, sat_prox_mat=S)
    assert env.scratch is not None  # float64 sums live in an L2-resident scratch buffer at this size
    batch = _batch_for(env, B, lazy=("beta", "avail_actions", "actions_onehot"))
    env.reset(batch)
    for t in range(T):
        env.step(th.tensor(acts[t], device="cuda"), batch)
    assert th.equal(batch["obs"].cpu(), _cast(want["obs"], th.float16))
    assert th.equal(batch["rewards"].cpu(), _cast(want["rewards"], th.float16))
    assert th.equal(batch["prev_assigns"].cpu(), _cast(want["prev_assigns"], th.int16))


@pytest.mark.parametrize("cfg", [
    dict(B=2, n=37, m=53, T=4, L=4, M=10, N=12, prios=False, dtype=th.float16, gen="ref", seed=21),    # keyed, ragged tiles
    dict(B=3, n=70, m=90, T=3, L=2, M=8, N=5, prios=True, dtype=th.float32, gen="dense", seed=22),     # keyed, priorities
    dict(B=2, n=130, m=140, T=3, L=3, M=12, N=10, prios=False, dtype=th.float16, gen="dense", seed=23),  # M + M/2 + 1 > 16: exact mode
    dict(B=2, n=64, m=96, T=3, L=5, M=10, N=10, prios=True, dtype=th.float32, gen="ref", seed=24),     # L > 4: exact mode
    dict(B=2, n=200, m=200, T=3, L=3, M=10, N=10, prios=False, dtype=th.float16, gen="ties", seed=25),  # natural large path, tie-heavy
    dict(B=1, n=511, m=511, T=2, L=2, M=10, N=15, prios=False, dtype=th.float16, gen="dense", seed=26),  # largest keyed shape
])
def test_real_env_multi_cta_path_matches_oracle(cfg, real_kernel_path):
    """The multi-CTA path of the large shapes (csrc/sap_real_large.cu) on ragged and extreme shapes, both of its modes."""
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    real_kernel_path(2)
    rng = np.random.default_rng(cfg["seed"])
    B, n, m, T, L, M, N = (cfg[k] for k in ("B", "n", "m", "T", "L", "M", "N"))
    if cfg["gen"] == "dense":
        S = O.gen_dense(rng, B, n, m, T)
    elif cfg["gen"] == "ties":
        S = (np.round(O.gen_exact(rng, B, n, m, T, zero_frac=0.5) * 4) / 4).astype(np.float32)
    else:
        S = O.gen_ref_like(rng, B, n, m, T)
    prios = (rng.integers(1, 4, size=m) * 0.5).astype(np.float32) if cfg["prios"] else None
    acts = rng.integers(0, m, size=(T, B, n))
    acts[:, :, : n // 3] = acts[:, :, :1]
    st = O.RealState(S.astype(np.float64), L, M, N, 0.5, task_prios=prios)
    want = O.rollout(st, lambda t, pre: acts[t], "real")
    env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5, sat_prox_mat=S, task_prios=prios)
    assert env.scratch is not None and env.launches_per_step == 4
    batch = _batch_for(env, B, cfg["dtype"])
    batch.agent_in = th.zeros(B, n, env.obs_size, device="cuda")
    env.reset(batch)
    counts = []
    for t in range(T):
        env.step(th.tensor(acts[t], device="cuda"), batch)
        counts.append(env.counts.cpu().numpy().copy())
        assert th.equal(batch.agent_in, batch["obs"][:, t + 1].float())
    td = {k: v.cpu() for k, v in batch.data.transition_data.items()}
    assert th.equal(td["obs"], _cast(want["obs"], cfg["dtype"]))
    assert th.equal(td["beta"], _cast(want["beta"], cfg["dtype"]))
    assert th.equal(td["rewards"], _cast(want["rewards"], cfg["dtype"]))
    assert th.equal(td["prev_assigns"], _cast(want["prev_assigns"], th.int16))
    assert th.equal(td["terminated"][..., 0], th.tensor(want["terminated"]))
    np.testing.assert_array_equal(np.stack(counts, 1), want["counts"][:, :T])
    np.testing.assert_allclose(env.ep_return.cpu().numpy(), want["rewards"].sum((1, 2)), rtol=1e-12)


def test_real_env_bench_batch_properties():
    """The bench configuration itself (4096 envs x 100 x 100): conflict histogram, sorted top-M, determinism."""
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    B, n, m, T, L, M, N = 4096, 100, 100, 3, 3, 10, 10
    g = th.Generator(device="cuda").manual_seed(5)
    planes = th.rand(B, T, n, m, device="cuda", generator=g)
    env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5)
    env.set_planes(planes)
    batch = _batch_for(env, B, lazy=("beta", "avail_actions", "actions_onehot"))
    acts = [th.randint(0, m, (B, n), device="cuda", generator=g) for _ in range(T)]
    sums = []
    for rep in range(2):
        env.reset(batch)
        for t in range(T):
            env.step(acts[t], batch)
            if rep == 0 and t + 1 < T:
                tot = planes[:, t + 1:t + 1 + L].double().sum(1)
                tv = tot.gather(2, env.top.long())
                assert bool((tv[..., :-1] >= tv[..., 1:]).all())
                assert bool((tot.scatter(2, env.top.long(), float("-inf")).max(-1).values <= tv[..., -1]).all())
                want_cnt = th.zeros(B, m, dtype=th.int32, device="cuda").scatter_add_(1, acts[t], th.ones_like(acts[t], dtype=th.int32))
                assert th.equal(env.counts, want_cnt)
        sums.append((batch["obs"].float().sum(dtype=th.float64).item(), batch["rewards"].float().sum(dtype=th.float64).item(),
                     batch["obs"].clone()))
    assert sums[0][0] == sums[1][0] and sums[0][1] == sums[1][1] and th.equal(sums[0][2], sums[1][2])


def test_benefit_generation_follows_the_reference_law():
    """sap_benefit_generate vs the law of generate_benefits_over_time (mock_constellation_env.py:276-299): per-task scale in
    {1, 10} (P = 3/4, 1/4), P(active) = 1/4, Gaussian bumps whose width lies in the requested range, reproducible per
    (seed, episode), and the same mean benefit as the oracle's numpy sampler of that law."""
    from marl_sap_b200.envs.batched import BatchedMockConstellationEnv

    B, n, m, T, L = 48, 20, 30, 60, 3
    env = BatchedMockConstellationEnv(B, n, m, T, L, 0.5, generate_seed=5)
    assert not env.constant_benefits
    ctor = env.planes.clone()           # widths 5..8 (:34)
    env.generate_benefits(3.0, 6.0)     # what reset() draws (:99-100)
    P = env.planes.double().cpu().numpy()  # [B, T, n, m]
    assert not np.array_equal(P, ctor.double().cpu().numpy())
    peak = P.max(1)                                     # [B, n, m]
    active = peak > 0
    assert abs(active.mean() - 0.25) < 0.02
    tmax = P.argmax(1)
    interior = active & (tmax > 8) & (tmax < T - 9)     # bumps whose centre is well inside the horizon
    logp = np.log(np.where(P > 0, P, 1.0))
    idx = np.argwhere(interior)
    b, i, j = idx[:, 0], idx[:, 1], idx[:, 2]
    t0 = tmax[b, i, j]
    # log of a Gaussian bump is a parabola: its second difference is -1 / sigma_2, and the centre gives the scale back
    d2 = logp[b, t0 + 1, i, j] - 2 * logp[b, t0, i, j] + logp[b, t0 - 1, i, j]
    sigma_2 = -1.0 / d2
    spread = np.sqrt(sigma_2 ** 2 * (-8 * np.log(0.05)))
    assert spread.min() > 3.0 - 1e-3 and spread.max() < 6.0 + 1e-3
    d1 = (logp[b, t0 + 1, i, j] - logp[b, t0 - 1, i, j]) / 2
    center = t0 + d1 * sigma_2
    scale = np.exp(logp[b, t0, i, j] + (t0 - center) ** 2 / sigma_2 / 2)
    assert np.all((np.abs(scale - 1) < 1e-3) | (np.abs(scale - 10) < 1e-2))
    per_task = np.round(scale).astype(int)
    for bb in range(4):  # one scale per (env, task), shared by all agents
        for jj in range(m):
            vals = per_task[(b == bb) & (j == jj)]
            assert len(set(vals.tolist())) <= 1
    assert abs((per_task == 10).mean() - 0.25) < 0.08
    # same law as the numpy sampler of the oracle
    ref = O.gen_ref_like(np.random.default_rng(0), B, n, m, T)  # [B, n, m, T]
    assert abs(P.mean() / ref.mean() - 1) < 0.15
    # reproducible: same (seed, episode) -> same planes; the episode counter advances at every reset
    env2 = BatchedMockConstellationEnv(B, n, m, T, L, 0.5, generate_seed=5)
    assert th.equal(env2.planes, ctor)
    batch = _batch_for(env2, B)
    env2.reset(batch)
    assert th.equal(env2.planes, env.planes)
    assert th.equal(batch["obs"][:, 0, :, m:2 * m].cpu(), env2.planes[:, 0].cpu())  # obs = [onehot | S[:, :, k] | ...]


def test_bids_as_actions_matches_reference_golden():
    """bids_as_actions through the single-env facades (reference API) and the batched envs: the bid matrix of every env
    becomes an assignment on the device, the buffer keeps the bids."""
    from marl_sap_b200.envs import REGISTRY
    from marl_sap_b200.envs.batched import BatchedMockConstellationEnv, BatchedRealConstellationEnv

    g = _load("bids.npz")
    S = g["S"].astype(np.float32)
    n, m, T = S.shape
    L, M, N, lam = int(g["L"]), int(g["M"]), int(g["N"]), float(g["lambda_"])
    env = REGISTRY["real_constellation_env"](num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=lam,
                                             sat_prox_mat=S, graphs=1, bids_as_actions=True)
    assert env.scheme["actions"]["vshape"] == (m,) and env.preprocess == {}
    env.reset()
    np.testing.assert_allclose(np.array(env.get_obs()), g["real_obs"][0], rtol=0, atol=0)
    for t, bids in enumerate(g["bids"]):
        r, d, _ = env.step(bids)
        np.testing.assert_allclose(r, g["real_rewards"][t], rtol=1e-6, atol=1e-7)  # facade reads rewards back at fp32
        np.testing.assert_allclose(np.array(env.get_obs()), g["real_obs"][t + 1], rtol=0, atol=0)
    np.testing.assert_array_equal(env.prev_assigns, g["real_prev"])
    # batched: B copies, bids stored in the batch
    B = 3
    benv = BatchedRealConstellationEnv(B, n, m, T, L, M, N, lam, sat_prox_mat=S)
    benv.enable_bids_as_actions()
    batch = _batch_for(benv, B, th.float32)
    assert batch["actions"].shape == (B, T + 1, n, m) and batch["actions"].dtype == th.float32
    benv.reset(batch)
    for t, bids in enumerate(g["bids"]):
        benv.step(th.tensor(np.broadcast_to(bids, (B, n, m)).copy(), device="cuda"), batch)
    assert th.equal(batch["rewards"][0, :T].cpu(), _cast(g["real_rewards"], th.float32))
    assert th.equal(batch["actions"][1, :T].cpu(), th.tensor(g["bids"]))
    menv = BatchedMockConstellationEnv(B, n, m, T, L, lam, sat_prox_mat=S)
    menv.enable_bids_as_actions()
    mb = _batch_for(menv, B)
    menv.reset(mb, prev0=np.broadcast_to(g["mock_prev0"], (B, n)).copy())
    for t, bids in enumerate(g["bids"]):
        menv.step(th.tensor(np.broadcast_to(bids, (B, n, m)).copy(), device="cuda"), mb)
    assert th.equal(mb["rewards"][2, :T].cpu(), _cast(g["mock_rewards"], th.float32))


@pytest.mark.parametrize("n,m,gen", [(50, 50, "dense"), (64, 100, "ties"), (33, 47, "ref")])
def test_real_env_small_common_config_matches_oracle(n, m, gen):
    """The bench's C2 shape family (n <= 64, fp16 scheme, M = N = 10, L = 3): the fast kernel's 4-lanes-per-list variant."""
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    rng = np.random.default_rng(n * 7 + m)
    B, T, L, M, N = 4, 5, 3, 10, 10
    if gen == "dense":
        S = O.gen_dense(rng, B, n, m, T)
    elif gen == "ties":
        S = (np.round(O.gen_exact(rng, B, n, m, T, zero_frac=0.5) * 4) / 4).astype(np.float32)
    else:
        S = O.gen_ref_like(rng, B, n, m, T)
    acts = rng.integers(0, m, size=(T, B, n))
    st = O.RealState(S.astype(np.float64), L, M, N, 0.5)
    want = O.rollout(st, lambda t, pre: acts[t], "real")
    env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5, sat_prox_mat=S)
    assert env.scratch is None
    batch = _batch_for(env, B, lazy=("beta", "avail_actions", "actions_onehot"))
    batch.agent_in = th.zeros(B, n, env.obs_size, device="cuda")
    env.reset(batch)
    for t in range(T):
        env.step(th.tensor(acts[t], device="cuda"), batch)
        assert th.equal(batch.agent_in, batch["obs"][:, t + 1].float())
    assert th.equal(batch["obs"].cpu(), _cast(want["obs"], th.float16))
    assert th.equal(batch["rewards"].cpu(), _cast(want["rewards"], th.float16))
    assert th.equal(batch["prev_assigns"].cpu(), _cast(want["prev_assigns"], th.int16))


def _gen_benefits(gen, rng, B, n, m, T):
    if gen == "dense":
        return O.gen_dense(rng, B, n, m, T)
    if gen == "ties":
        return (np.round(O.gen_exact(rng, B, n, m, T, zero_frac=0.5) * 4) / 4).astype(np.float32)
    if gen == "dup":      # duplicates above the minimum + near-ties one ulp apart: exercises the exact redo path
        S = (np.round(O.gen_dense(rng, B, n, m, T) * 16) / 16 + 1).astype(np.float32)
        S[:, ::3] = np.nextafter(S[:, ::3], np.float32(4))
        return S
    if gen == "neg":      # negative benefits: every key is inexact, every list goes through the exact selection
        return (O.gen_ref_like(rng, B, n, m, T) - np.float32(0.25) * (rng.random((B, n, m, T)) < 0.1)).astype(np.float32)
    if gen == "const":    # every window sum identical
        return np.full((B, n, m, T), 0.5, dtype=np.float32)
    return O.gen_ref_like(rng, B, n, m, T)


@pytest.mark.parametrize("n,m,gen,opts", [
    (100, 100, "dense", ""), (100, 100, "ties", ""), (100, 100, "ref", "eager"), (100, 100, "dup", ""),
    (100, 100, "neg", ""), (100, 100, "const", ""), (68, 72, "dense", "noain"), (128, 128, "ties", ""),
    (96, 128, "ref", "ttrans"), (72, 100, "dup", "shared"), (124, 128, "dense", "eager"),
])
def test_real_env_fast2_matches_oracle(n, m, gen, opts):
    """The bench shape family (64 < n <= 128, fp16 scheme, M = N = 10, L = 3): the second-generation one-CTA-per-env
    kernel (csrc/sap_real_fast2.cu: transposed key tile, slot gather) against the oracle, including tie-heavy,
    near-tie, negative and constant benefits, eager beta / avail / onehot fields, a custom T_trans and shared planes."""
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    rng = np.random.default_rng(n * 13 + m + len(gen))
    B, T, L, M, N = 3, 5, 3, 10, 10
    shared = opts == "shared"
    S = _gen_benefits(gen, rng, 1 if shared else B, n, m, T)
    Tt = rng.integers(0, 2, size=(m, m)).astype(np.float64) if opts == "ttrans" else None
    acts = rng.integers(0, m, size=(T, B, n))
    acts[:, :, : n // 3] = acts[:, :, :1]
    S_or = np.broadcast_to(S, (B, n, m, T)) if shared else S
    st = O.RealState(S_or.astype(np.float64), L, M, N, 0.5, T_trans=Tt)
    want = O.rollout(st, lambda t, pre: acts[t], "real")
    env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5, sat_prox_mat=S[0] if shared else S, T_trans=Tt)
    assert env.scratch is None
    batch = _batch_for(env, B, lazy=() if opts == "eager" else ("beta", "avail_actions", "actions_onehot"))
    if opts != "noain":
        batch.agent_in = th.zeros(B, n, env.obs_size, device="cuda")
    env.reset(batch)
    tops, counts = [env.top.cpu().numpy().copy()], []
    for t in range(T):
        env.step(th.tensor(acts[t], device="cuda"), batch)
        tops.append(env.top.cpu().numpy().copy())
        counts.append(env.counts.cpu().numpy().copy())
        if opts != "noain":
            assert th.equal(batch.agent_in, batch["obs"][:, t + 1].float())
    assert th.equal(batch["obs"].cpu(), _cast(want["obs"], th.float16))
    assert th.equal(batch["rewards"].cpu(), _cast(want["rewards"], th.float16))
    assert th.equal(batch["prev_assigns"].cpu(), _cast(want["prev_assigns"], th.int16))
    assert th.equal(batch["terminated"][..., 0].cpu(), th.tensor(want["terminated"]))
    assert th.equal(batch["filled"][..., 0].cpu(), th.tensor(want["filled"]))
    np.testing.assert_array_equal(np.stack(counts, 1), want["counts"][:, :T])
    for t in range(T):
        np.testing.assert_array_equal(tops[t], O.top_m_tasks(want["beta"][:, t], M))
    np.testing.assert_allclose(env.ep_return.cpu().numpy(), want["rewards"].sum((1, 2)), rtol=1e-12)
    if opts == "eager":
        assert th.equal(batch["beta"].cpu(), _cast(want["beta"], th.float16))
        assert bool(batch["avail_actions"].all())
        assert th.equal(batch["actions_onehot"][:, :T].cpu(), _cast(O.one_hot(np.moveaxis(acts, 0, 1), m, np.int16), th.int16))


@pytest.mark.parametrize("name", ["power_env.npz", "interference_env.npz"])
def test_power_and_interference_envs_match_reference_golden(name):
    """RealPowerConstellationEnv / InterferenceConstellationEnv through the batched device envs and the
    reference-API facades, against vectors recorded from the unmodified reference classes: observations with the N + 1
    power columns, rewards (zero / interference rule for agents out of power), prev_assigns, the power_states field and the
    float64 power trajectory itself (the 5.55e-17 residue and the -0.2 it becomes)."""
    from marl_sap_b200.envs import REGISTRY
    from marl_sap_b200.envs.batched import BatchedInterferenceConstellationEnv, BatchedRealPowerConstellationEnv

    g = _load(name)
    S = g["S"]
    n, m, T = S.shape
    L, M, N, lam = int(g["L"]), int(g["M"]), int(g["N"]), float(g["lambda_"])
    B = 3
    if name.startswith("power"):
        env = BatchedRealPowerConstellationEnv(B, n, m, T, L, M, N, lam, sat_prox_mat=S, task_prios=g["task_prios"])
    else:
        env = BatchedInterferenceConstellationEnv(B, n, m, T, L, M, N, lam, S, g["neighbor_matrix"], g["sat_freq_bands"],
                                                  task_prios=g["task_prios"])
    assert env.obs_size == int(g["obs_size"])
    batch = _batch_for(env, B)
    batch.agent_in = th.zeros(B, n, env.obs_size, device="cuda")
    env.reset(batch, prev0=np.broadcast_to(g["prev0"], (B, n)).copy())
    acts = g["actions"]
    powers = [env.power.cpu().numpy().copy()]
    for t in range(T):
        done = env.step(th.tensor(np.broadcast_to(acts[t], (B, n)).copy(), device="cuda"), batch)
        assert done == bool(g["done"][t])
        powers.append(env.power.cpu().numpy().copy())
        assert th.equal(batch.agent_in, batch["obs"][:, t + 1].float())
    td = {k: v.cpu() for k, v in batch.data.transition_data.items()}
    for b in range(B):
        assert th.equal(td["obs"][b], _cast(g["obs"], th.float16)), f"obs mismatch env {b}"
        assert th.equal(td["beta"][b], _cast(g["beta"], th.float16))
        assert th.equal(td["prev_assigns"][b], _cast(g["prev"], th.int16))
        assert th.equal(td["rewards"][b, :T], _cast(g["rewards"], th.float16))
        assert th.equal(td["power_states"][b], _cast(g["power"], th.float16))
        assert td["terminated"][b, :T, 0].tolist() == [bool(d) for d in g["done"]]
        np.testing.assert_array_equal(np.stack(powers)[:, b], g["power"])            # float64, bit for bit
    np.testing.assert_allclose(env.ep_return.cpu().numpy(), np.full(B, g["rewards"].sum()), rtol=1e-12)
    # the reference-API facade (REGISTRY object driven like the reference env)
    np.random.seed(0)
    if name.startswith("power"):
        fac = REGISTRY["real_power_constellation_env"](num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=lam,
                                                       sat_prox_mat=S.astype(np.float64), graphs=1, task_prios=g["task_prios"])
    else:
        fac = REGISTRY["interference_constellation_env"](num_planes=1, num_sats_per_plane=n, res=2, T=T, N=N, M=M, L=L,
                                                         lambda_=lam, task_prios=g["task_prios"],
                                                         sat_freq_bands=g["sat_freq_bands"], sat_prox_mat=S.astype(np.float64),
                                                         neighbor_matrix=g["neighbor_matrix"])
    assert fac.scheme["power_states"]["vshape"] == (n,) and fac.get_obs_size() == int(g["obs_size"])
    fac.reset()
    pre = fac.get_pretransition_data()
    assert set(pre) == {"beta", "obs", "prev_assigns", "avail_actions", "power_states"}
    np.testing.assert_array_equal(pre["power_states"][0], np.ones(n))
    r, d, info = fac.step(list(acts[0]))
    assert len(r) == n and d is False and info == {}
    bh = fac.beta_hat(pre["beta"][0], pre["prev_assigns"][0], np.array([0.0] + [1.0] * (n - 1)))
    assert not bh[0].any() and bh[1:].any()


@pytest.mark.parametrize("kind", ["power", "interference"])
def test_power_and_interference_envs_match_oracle(kind):
    """Seeded batches (distinct benefits per env, priorities, conflicts, agents that die and recharge) against the oracle."""
    from marl_sap_b200.envs.batched import BatchedInterferenceConstellationEnv, BatchedRealPowerConstellationEnv

    rng = np.random.default_rng(41)
    B, n, m, T, L, M, N, lam = 4, 20, 30, 12, 3, 10, 10, 0.5
    S = O.gen_ref_like(rng, B, n, m, T) + O.gen_dense(rng, B, n, m, T) * (rng.random((B, n, m, 1)) < 0.5)
    S = S.astype(np.float32)
    prios = rng.choice([1.0, 1.0, 1.0, 5.0], size=m)
    acts = rng.integers(0, m, size=(T, B, n))
    acts[:, :, : n // 3] = acts[:, :, :1]
    acts[:, :, n // 3] = np.argmax(S.sum(-1)[:, n // 3], axis=-1)[None]   # an agent that keeps working its best task
    prev0 = np.stack([rng.permutation(m)[:n] for _ in range(B)])
    if kind == "power":
        st = O.PowerState(S.astype(np.float64), L, M, N, lam, task_prios=prios)
        env = BatchedRealPowerConstellationEnv(B, n, m, T, L, M, N, lam, sat_prox_mat=S, task_prios=prios)
    else:
        nb = (rng.random((m, m)) < 0.2).astype(np.float64)
        nb = np.maximum(nb, nb.T)
        np.fill_diagonal(nb, 1.0)
        bands = rng.integers(0, 4, size=(B, n))
        st = O.InterferenceState(S.astype(np.float64), L, M, N, lam, nb, bands, task_prios=prios)
        env = BatchedInterferenceConstellationEnv(B, n, m, T, L, M, N, lam, S, nb, bands, task_prios=prios)
    want = O.rollout(st, lambda t, pre: acts[t], "real", prev0=prev0)
    batch = _batch_for(env, B, th.float32)
    env.reset(batch, prev0=prev0)
    for t in range(T):
        env.step(th.tensor(acts[t], device="cuda"), batch)
    td = {k: v.cpu() for k, v in batch.data.transition_data.items()}
    assert th.equal(td["obs"], _cast(want["obs"], th.float32))
    assert th.equal(td["rewards"], _cast(want["rewards"], th.float32))
    assert th.equal(td["prev_assigns"], _cast(want["prev_assigns"], th.int16))
    assert th.equal(td["power_states"], _cast(want["power_states"], th.float16))
    np.testing.assert_array_equal(env.power.cpu().numpy(), want["power_states"][:, T])
    np.testing.assert_allclose(env.ep_return.cpu().numpy(), want["rewards"].sum((1, 2)), rtol=1e-12)
    assert (want["power_states"] <= 0).any()


def test_fov_proximities_match_the_reference_function():
    """sap_proximities_fov vs calc_fov_based_proximities_fast of the reference (HighPerformanceConstellationSim.py:308-327,
    golden tests/golden/proximities.npz): the visible / invisible pattern exactly, the values to 1e-12 in float64 (acos / exp
    of CUDA vs numpy's libm) and to fp32 rounding in the plane layout the env kernels read; then an env runs on them."""
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv
    from marl_sap_b200.envs.proximity import fov_proximities, gaussian_sigma_2

    g = _load("proximities.npz")
    assert gaussian_sigma_2(float(g["fov"])) == pytest.approx(float(g["sigma_2"]), rel=1e-15)
    planes, ref = fov_proximities(g["sat_r"], g["task_r"], fov=float(g["fov"]), reference_layout=True)
    want = g["prox"]
    got = ref.cpu().numpy()
    np.testing.assert_array_equal(got > 0, want > 0)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    np.testing.assert_allclose(planes.cpu().numpy(), np.transpose(want, (2, 0, 1)).astype(np.float32), rtol=2e-7, atol=0)
    n, m, T = want.shape
    env = BatchedRealConstellationEnv(2, n, m, T, 3, 4, 3, 0.5)
    env.set_planes(planes[None], shared=True)
    batch = _batch_for(env, 2)
    env.reset(batch)
    st = O.RealState(np.broadcast_to(planes.cpu().numpy().transpose(1, 2, 0).astype(np.float64), (2, n, m, T)).copy(), 3, 4, 3, 0.5)
    st.reset()
    assert th.equal(batch["obs"][:, 0].cpu(), _cast(st.obs, th.float16))
