"""GPU parity of the selector and episode-buffer kernels (through the C ABI)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch as th

from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _args(**kw):
    base = dict(epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=1000, evaluation_epsilon=0.0,
                env_args={"M": 4, "m": 9}, seed=7)
    base.update(kw)
    return SimpleNamespace(**base)


def _cu(x, dtype=None):
    t = th.tensor(np.asarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def test_selectors_match_reference_golden():
    from marl_sap_b200.action_selectors import REGISTRY

    g = dict(np.load(os.path.join(GOLDEN, "selectors.npz")))
    sel = REGISTRY["epsilon_greedy"](_args())
    t_envs = list(g["eg_t_env"]) + ["test"]
    for t_env, want in zip(t_envs, g["eg_actions"]):
        sel.inject_draws(u_explore=_cu(g["eg_u_explore"]), u_action=_cu(g["eg_u_action"]))
        if t_env == "test":
            got = sel.select_action(_cu(g["eg_q"]), _cu(g["eg_avail"]), 0, test_mode=True)
        else:
            got = sel.select_action(_cu(g["eg_q"]), _cu(g["eg_avail"]), int(t_env), test_mode=False)
        assert got.dtype == th.int64
        np.testing.assert_array_equal(got.cpu().numpy(), want)
    assert sel.epsilon == 0.0
    fsel = REGISTRY["filtered_const_epsilon_greedy"](_args())
    for t_env, want in zip(t_envs, g["fg_actions"]):
        fsel.inject_draws(u_tie=_cu(g["fg_u_tie"]), u_explore=_cu(g["fg_u_explore"]), u_action=_cu(g["fg_u_action"]))
        kw = dict(test_mode=True) if t_env == "test" else dict(test_mode=False)
        got = fsel.select_action(_cu(g["fg_q"]), _cu(g["eg_avail"]), 0 if t_env == "test" else int(t_env),
                                 beta=_cu(g["fg_beta"]), **kw)
        np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_sap_selectors_match_reference_golden():
    """The four assignment selectors against the picks of the unmodified reference (scipy per env), same draws."""
    from marl_sap_b200.action_selectors import REGISTRY

    g = dict(np.load(os.path.join(GOLDEN, "sap_selectors.npz")))
    M, m = int(g["sap_M"]), int(g["sap_m"])
    args = _args(env_args={"M": M, "m": m})
    t_envs = list(g["sap_t_env"]) + ["test"]
    avail = th.ones(g["sap_q"].shape, dtype=th.bool, device="cuda")
    sel = REGISTRY["sap"](args)
    for t_env, want in zip(t_envs, g["sap_actions"]):
        sel.inject_draws(z=_cu(g["sap_z"]))
        got = sel.select_action(_cu(g["sap_q"]), avail, 0 if t_env == "test" else int(t_env), test_mode=t_env == "test")
        assert got.dtype == th.int64 and got.is_cuda
        np.testing.assert_array_equal(got.cpu().numpy(), want)
    got = REGISTRY["epsilon_greedy_sap_test"](args).select_action(_cu(g["sap_q"]), avail, 0, test_mode=True)
    np.testing.assert_array_equal(got.cpu().numpy(), g["egsap_test_actions"])
    fsel = REGISTRY["filtered_const_sap"](args)
    for t_env, want in zip(t_envs, g["fsap_actions"]):
        fsel.inject_draws(z=_cu(g["fsap_z"]), u_tie=_cu(g["fsap_u_tie"]))
        got = fsel.select_action(_cu(g["fsap_q"]), avail, 0 if t_env == "test" else int(t_env), test_mode=t_env == "test",
                                 beta=_cu(g["fsap_beta"]))
        np.testing.assert_array_equal(got.cpu().numpy(), want)
    fsel2 = REGISTRY["filtered_const_epsgr_sap_test"](args)
    fsel2.inject_draws(u_tie=_cu(g["fsap_u_tie"]))
    got = fsel2.select_action(_cu(g["fsap_q"]), avail, 0, test_mode=True, beta=_cu(g["fsap_beta"]))
    np.testing.assert_array_equal(got.cpu().numpy(), g["fepsgr_test_actions"])
    # training branches delegate to the epsilon-greedy kernels: actions stay in range, epsilon follows the schedule
    a = fsel2.select_action(_cu(g["fsap_q"]), avail, 500, test_mode=False, beta=_cu(g["fsap_beta"]))
    assert a.shape == (g["fsap_q"].shape[0], g["fsap_q"].shape[1]) and int(a.min()) >= 0 and int(a.max()) < m


def test_policy_selectors_match_reference_golden():
    """multinomial / soft_policies / filtered_const_soft_policies against the reference's picks (same uniforms)."""
    from marl_sap_b200.action_selectors import REGISTRY

    g = dict(np.load(os.path.join(GOLDEN, "policy_selectors.npz")))
    M, m = int(g["M"]), g["p"].shape[2]
    args = _args(env_args={"M": M, "m": m}, test_greedy=True)
    msel = REGISTRY["multinomial"](args)
    msel.inject_draws(u_sample=_cu(g["u"]))
    np.testing.assert_array_equal(msel.select_action(_cu(g["p"]), _cu(g["avail"]), 0).cpu().numpy(), g["multinomial_train"])
    np.testing.assert_array_equal(msel.select_action(_cu(g["p"]), _cu(g["avail"]), 0, test_mode=True).cpu().numpy(),
                                  g["multinomial_test"])
    ssel = REGISTRY["soft_policies"](args)
    ssel.inject_draws(u_sample=_cu(g["u"]))
    np.testing.assert_array_equal(ssel.select_action(_cu(g["p"]), _cu(g["avail"]), 0).cpu().numpy(), g["soft"])
    fsel = REGISTRY["filtered_const_soft_policies"](args)
    fsel.inject_draws(u_sample=_cu(g["uf"]), u_rand=_cu(g["u_rand"]))
    got = fsel.select_action(_cu(g["pf"]), _cu(g["avail"]), 0, beta=_cu(g["beta"]))
    np.testing.assert_array_equal(got.cpu().numpy(), g["filtered"])


@pytest.mark.parametrize("rows,A,masked", [(1000, 100, True), (77, 450, False), (33, 5, True), (4, 1, False)])
def test_sample_categorical_matches_oracle(rows, A, masked):
    """sap_sample_categorical vs the float64 inverse-CDF contract; the empirical distribution follows the probabilities."""
    from marl_sap_b200.action_selectors.policy_selectors import sample_categorical

    rng = np.random.default_rng(rows + A)
    p = rng.random((rows, A)).astype(np.float32)
    p[:, ::3] = 0.0 if A > 3 else p[:, ::3]
    avail = (rng.random((rows, A)) > 0.3) if masked else None
    if masked:
        avail[:, -1] = True
    u = rng.random(rows, dtype=np.float32)
    u[:2] = [0.0, np.float32(1.0 - 2.0 ** -24)][: min(2, rows)]
    got = sample_categorical(_cu(p), None if avail is None else _cu(avail), _cu(u)).cpu().numpy()
    np.testing.assert_array_equal(got, O.sample_categorical(p, u, avail))
    eff = p if avail is None else p * avail
    assert (eff[np.arange(rows), got] > 0).all()  # a zero-probability action is never drawn
    if rows >= 1000:  # the kernel's own uniforms: frequencies of one fixed distribution
        q = np.tile(np.array([0.5, 0.0, 0.3, 0.2], np.float32), (20000, 1))
        freq = np.bincount(sample_categorical(_cu(q)).cpu().numpy(), minlength=4) / 20000
        np.testing.assert_allclose(freq, [0.5, 0.0, 0.3, 0.2], atol=0.02)


@pytest.mark.parametrize("B,n,m,noise", [(8, 4, 4, False), (33, 10, 10, True), (16, 50, 50, True), (64, 100, 100, True),
                                         (5, 37, 53, True), (3, 324, 450, True), (2, 200, 512, False), (4, 1, 7, True)])
def test_lsa_kernel_matches_scipy(B, n, m, noise):
    """sap_lsa_maximize against scipy.optimize.linear_sum_assignment (through the oracle): feasible, same optimal
    objective to 1e-9 relative, and - the optimum being unique for continuous inputs - the same assignment."""
    from marl_sap_b200.action_selectors.sap_selectors import lsa_maximize

    rng = np.random.default_rng(B * 1000 + m)
    q = rng.standard_normal((B, n, m)).astype(np.float32)
    z = rng.standard_normal((B, n, m)).astype(np.float32) if noise else None
    std = O.sap_noise_std(q, 0.3) if noise else None
    want, want_obj = O.lsa_maximize(q, z, std)
    got, obj = lsa_maximize(_cu(q), None if z is None else _cu(z), None if std is None else _cu(std), want_objective=True)
    got, obj = got.cpu().numpy(), obj.cpu().numpy()
    for b in range(B):
        assert len(set(got[b].tolist())) == n and got[b].min() >= 0 and got[b].max() < m
    np.testing.assert_allclose(obj, want_obj, rtol=1e-9, atol=1e-9)
    np.testing.assert_array_equal(got, want)


def test_lsa_kernel_ties_and_errors():
    """Exact ties (constant and integer matrices): any optimal assignment is accepted, the objective must be optimal."""
    from marl_sap_b200 import _lib
    from marl_sap_b200.action_selectors.sap_selectors import lsa_maximize

    rng = np.random.default_rng(5)
    q = np.stack([np.zeros((6, 9), np.float32), np.ones((6, 9), np.float32),
                  rng.integers(0, 3, size=(6, 9)).astype(np.float32), rng.integers(-2, 2, size=(6, 9)).astype(np.float32)])
    want, want_obj = O.lsa_maximize(q)
    got, obj = lsa_maximize(_cu(q), want_objective=True)
    got = got.cpu().numpy()
    for b in range(q.shape[0]):
        assert len(set(got[b].tolist())) == 6
    np.testing.assert_allclose(obj.cpu().numpy(), want_obj, rtol=0, atol=1e-12)
    with pytest.raises(RuntimeError, match="n <= m"):
        lsa_maximize(_cu(np.zeros((1, 5, 3), np.float32)))
    assert _lib.load().sap_lsa_maximize(None, None, None, 1, 1, 1, None, None, None) != 0


@pytest.mark.parametrize("B,n,A", [(4, 10, 10), (3, 50, 50), (2, 100, 100), (2, 7, 450), (1, 3, 33)])
def test_epsilon_greedy_matches_oracle(B, n, A):
    from marl_sap_b200.action_selectors import REGISTRY

    rng = np.random.default_rng(B * 100 + A)
    q = rng.standard_normal((B, n, A)).astype(np.float32)
    q[0, 0, :] = 1.0  # full tie -> index 0
    q[0, 1, A // 2:] = 5.0  # tie -> first of the tied block
    avail = rng.random((B, n, A)) > 0.4
    avail[..., -1] |= ~avail.any(-1)
    avail[-1, -1, :] = False
    avail[-1, -1, A - 1] = True  # single available action
    ue = rng.random((B, n), dtype=np.float32)
    ua = rng.random((B, n), dtype=np.float32)
    ua[0, 0] = np.float32(1.0) - np.float32(2.0 ** -24)  # largest fp32 below 1
    sel = REGISTRY["epsilon_greedy"](_args())
    for t_env in (0, 400, 990, 5000):
        eps = O.epsilon_linear(1.0, 0.05, 1000, t_env)
        want = O.select_epsilon_greedy(q, avail, eps, ue, ua)
        sel.inject_draws(u_explore=_cu(ue), u_action=_cu(ua))
        got = sel.select_action(_cu(q), _cu(avail), t_env)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
        assert sel.epsilon == eps
    # no mask given (lazy all-ones avail): same as an all-True mask
    want = O.select_epsilon_greedy(q, np.ones_like(avail), 0.3, ue, ua)
    sel2 = REGISTRY["epsilon_greedy"](_args(epsilon_start=0.3, epsilon_finish=0.3))
    sel2.inject_draws(u_explore=_cu(ue), u_action=_cu(ua))
    got = sel2.select_action(_cu(q), th.ones(1, dtype=th.bool, device="cuda").expand(B, n, A), 0)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


@pytest.mark.parametrize("B,n,m,M,L", [(3, 10, 24, 6, 3), (2, 50, 50, 10, 3), (2, 20, 450, 10, 3)])
def test_filtered_epsilon_greedy_matches_oracle(B, n, m, M, L):
    from marl_sap_b200.action_selectors import REGISTRY

    rng = np.random.default_rng(m)
    q = rng.standard_normal((B, n, M + 1)).astype(np.float32)
    q[0, 0, :M] = -9.0  # baseline wins: decided by the 1e-8 tie noise
    q[0, 1, :M] = -9.0
    q[0, 1, M] = 0.75  # |base| >= 0.25: noise vanishes in fp32 -> first non-top index
    beta = O.gen_exact(rng, B, n, m, L)
    beta[1, 0, :, :] = 0.25  # all rows tie -> stable top-M = 0..M-1
    avail = rng.random((B, n, m)) > 0.2
    avail[..., 0] = True
    ut = rng.random((B, n, m), dtype=np.float32)
    ue = rng.random((B, n), dtype=np.float32)
    ua = rng.random((B, n), dtype=np.float32)
    top = O.top_m_tasks(beta, M)
    sel = REGISTRY["filtered_const_epsilon_greedy"](_args(env_args={"M": M, "m": m}))
    for t_env in (0, 600, 5000):
        eps = O.epsilon_linear(1.0, 0.05, 1000, t_env)
        want = O.select_filtered_epsilon_greedy(q, top, avail, m, eps, ut, ue, ua)
        for use_top in (False, True):
            sel.inject_draws(u_tie=_cu(ut), u_explore=_cu(ue), u_action=_cu(ua))
            if use_top:
                got = sel.select_action(_cu(q), _cu(avail), t_env, top=_cu(top, th.int32))
            else:
                got = sel.select_action(_cu(q), _cu(avail), t_env, beta=_cu(beta))
            np.testing.assert_array_equal(got.cpu().numpy(), want)
    # fp16 beta (the real scheme's buffer dtype) with exactly representable values gives the same top-M
    sel.inject_draws(u_tie=_cu(ut), u_explore=_cu(ue), u_action=_cu(ua))
    b16 = (np.round(beta * 64) / 64).astype(np.float16)
    want = O.select_filtered_epsilon_greedy(q, O.top_m_tasks(b16.astype(np.float64), M), avail, m, 0.05, ut, ue, ua)
    got = sel.select_action(_cu(q), _cu(avail), 5000, beta=_cu(b16))
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_philox_selection_statistics_and_determinism():
    """In-kernel RNG: explore rate ~ eps, random actions uniform over the available set, reproducible."""
    from marl_sap_b200.action_selectors import REGISTRY

    B, n, A = 256, 100, 20
    q = th.zeros(B, n, A, device="cuda")
    q[..., 3] = 1.0  # greedy action = 3
    avail = th.ones(B, n, A, dtype=th.bool, device="cuda")
    avail[..., 10:] = False
    avail[..., 3] = True
    sel = REGISTRY["epsilon_greedy"](_args(epsilon_start=0.5, epsilon_finish=0.5))
    ctr = th.zeros(1, dtype=th.int64, device="cuda")
    k = th.zeros(B, dtype=th.int32, device="cuda")
    sel.bind_counters(ctr, k)
    a0 = sel.select_action(q, avail, 0)
    a0b = sel.select_action(q, avail, 0)
    assert th.equal(a0, a0b)  # same (seed, episode, step) -> same draws
    k += 1
    a1 = sel.select_action(q, avail, 0)
    assert not th.equal(a0, a1)
    ctr += 1
    k.zero_()
    a2 = sel.select_action(q, avail, 0)
    assert not th.equal(a0, a2)
    assert bool(avail.gather(2, a0[..., None]).all())
    # explore w.p. 0.5, and an explored action is 3 w.p. 1/10 -> P(a != 3) = 0.45
    frac = (a0 != 3).float().mean().item()
    assert abs(frac - 0.45) < 0.01
    hist = th.bincount(a0[a0 != 3].flatten(), minlength=A).float()
    assert hist[10:].sum() == 0
    expected = hist.sum() / 9
    assert bool(((hist[:10][th.arange(10) != 3] - expected).abs() < 5 * expected.sqrt()).all())
    other = REGISTRY["epsilon_greedy"](_args(epsilon_start=0.5, epsilon_finish=0.5, seed=8))
    other.bind_counters(ctr, k)
    assert not th.equal(other.select_action(q, avail, 0), a2)


def test_benefit_ingest_and_beta_window():
    from marl_sap_b200 import _lib
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    B, n, m, T, L = 3, 7, 45, 37, 3
    S = th.rand(B, n, m, T)
    env = BatchedRealConstellationEnv(B, n, m, T, L, 4, 3, 0.5)
    env.load_benefits(S.numpy())  # host path: upload + re-layout
    want = S.permute(0, 3, 1, 2).contiguous()
    assert th.equal(env.planes.cpu(), want)
    env.load_benefits(S.cuda())  # device path
    assert th.equal(env.planes.cpu(), want)
    env.load_benefits(S[1])  # shared [n,m,T]
    assert env.planes.shape[0] == 1 and th.equal(env.planes.cpu()[0], want[1])
    env.load_benefits(S.numpy())
    beta = env.beta_field(th.float32).cpu()
    wb = th.zeros(B, T + 1, n, m, L)
    for t in range(T):
        for l in range(L):
            if t + l < T:
                wb[:, t, :, :, l] = S[..., t + l]
    assert th.equal(beta, wb)
    assert _lib.load().sap_abi_version() == 1


def test_replay_buffer_matches_reference_golden():
    """ReplayBuffer ring insert + EpisodeBatch.update casting vs the reference (tests/golden/buffer.npz)."""
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
    rb = ReplayBuffer(dict(scheme), groups, 5, T + 1, preprocess=pre, device="cuda")
    for e in range(4):
        B = 2
        batch = EpisodeBatch(dict(scheme), groups, B, T + 1, preprocess=pre, device="cuda")
        for t in range(T + 1):
            batch.update({"obs": [g[f"raw{e}_obs"][b, t] for b in range(B)],
                          "beta": [g[f"raw{e}_beta"][b, t] for b in range(B)],
                          "avail_actions": [[[1] * m] * n for b in range(B)],
                          "prev_assigns": [np.arange(n) for b in range(B)]}, ts=t)
            if t < T:
                batch.update({"actions": th.tensor(g[f"raw{e}_actions"][:, t]).cuda(),
                              "rewards": [(list(g[f"raw{e}_rewards"][b, t]),) for b in range(B)],
                              "terminated": [(t == T - 1,) for b in range(B)]}, ts=t)
        for k, v in batch.data.transition_data.items():
            np.testing.assert_array_equal(v.cpu().numpy(), g[f"ep{e}_{k}"], err_msg=f"episode {e} field {k}")
        rb.insert_episode_batch(batch)  # 4 x 2 episodes into a ring of 5 -> wraps
    for k, v in rb.data.transition_data.items():
        np.testing.assert_array_equal(v.cpu().numpy(), g[f"rb_{k}"], err_msg=f"replay field {k}")
    assert rb.buffer_index == int(g["rb_buffer_index"]) and rb.episodes_in_buffer == int(g["rb_episodes_in_buffer"])
    assert int(rb.max_t_filled()) == int(g["max_t_filled"])
    # sample(): a copy of the chosen episodes
    np.random.seed(0)
    ids = np.random.choice(rb.episodes_in_buffer, 3, replace=False)
    np.random.seed(0)
    smp = rb.sample(3)
    for k, v in smp.data.transition_data.items():
        np.testing.assert_array_equal(v.cpu().numpy(), g[f"rb_{k}"][ids])
    full = rb.sample(rb.episodes_in_buffer)
    assert full.batch_size == rb.episodes_in_buffer
    sub = rb[1:3, :T]
    assert sub.batch_size == 2 and sub.max_seq_length == T and sub["obs"].shape[:2] == (2, T)
    with pytest.raises(KeyError):
        batch.update({"nope": [1]}, ts=0)


def test_onehot_kernel_dtypes():
    from marl_sap_b200.components.transforms import OneHot

    a = th.randint(0, 11, (5, 7, 1), device="cuda")
    for dt in (th.int64, th.int16, th.int32):
        oh = OneHot(11).transform(a.to(dt))
        assert oh.dtype == th.float32
        assert th.equal(oh, th.nn.functional.one_hot(a[..., 0], 11).float())


def test_lsa_kernel_terminates_on_nan_and_inf_input():
    """Non-finite Q-values must not hang the assignment kernel (every loop is bounded); finite envs are unaffected."""
    from marl_sap_b200.action_selectors.sap_selectors import lsa_maximize

    rng = np.random.default_rng(2)
    q = rng.standard_normal((6, 12, 12)).astype(np.float32)
    q[1, 3, :] = np.nan
    q[2] = np.nan
    q[3, :, 5] = -np.inf
    q[4, 0, 0] = np.inf
    got = lsa_maximize(_cu(q)).cpu().numpy()
    want, _ = O.lsa_maximize(q[[0, 5]])
    np.testing.assert_array_equal(got[[0, 5]], want)
    assert got.shape == (6, 12)
