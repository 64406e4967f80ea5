"""The CPU oracle replayed against vectors recorded from the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle (prompt section 3)."""
import glob
import os

import numpy as np
import pytest

from oracle import cpu_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REAL = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "real_*.npz"))) + ["kat1_real.npz"]
MOCK = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "mock_*.npz")))


def _load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def test_kat1_known_answers():
    """KAT-1 values as listed in SURVEY.md section 8c (author's fixture, experiments.py:265-288)."""
    g = _load("kat1_real.npz")
    assert (int(g["n"]), int(g["m"]), int(g["T"]), int(g["L"])) == (4, 4, 2, 1)
    np.testing.assert_array_equal(g["obs"][0][0], [5, 1, 1, 10, 3, 2, 3, 4, 1, 0])
    np.testing.assert_array_equal(g["obs"][0][3], [10, 3, 2, 1, 1, 0, 4, 5, 1, 0])
    np.testing.assert_allclose(g["rewards"][0], [2.5, 0.75, 4.0, 10.0])
    np.testing.assert_allclose(g["rewards"][1], [1 / 6, 1 / 6, 1 / 6, 1.0])
    assert list(g["done"]) == [False, True]
    assert not g["obs"][2].any() and not g["beta"][2].any()


@pytest.mark.parametrize("name", REAL)
def test_real_env_oracle_matches_reference(name):
    g = _load(name)
    S = g["S"].astype(np.float64)[None]
    st = O.RealState(S, int(g["L_arg"]), int(g["M"]), int(g["N"]), float(g["lambda_"]),
                     task_prios=g.get("task_prios"), T_ctor=int(g["T_ctor"]))
    assert (st.n, st.m, st.T, st.L, st.obs_size) == tuple(int(g[k]) for k in ("n", "m", "T", "L", "obs_size"))
    st.reset()
    for t in range(g["actions"].shape[0] + 1):
        pre = st.pretransition()
        np.testing.assert_array_equal(pre["obs"][0], g["obs"][t])          # bit-exact: values are gathers
        np.testing.assert_array_equal(pre["beta"][0], g["beta"][t])
        np.testing.assert_array_equal(pre["prev_assigns"][0], g["prev"][t])
        if t < g["actions"].shape[0]:
            r, d = st.step(g["actions"][t][None])
            np.testing.assert_array_equal(r[0], g["rewards"][t])           # same float64 ops -> identical
            assert d == bool(g["done"][t])


@pytest.mark.parametrize("name", MOCK)
def test_mock_env_oracle_matches_reference(name):
    g = _load(name)
    st = O.MockState(g["S"].astype(np.float64)[None], int(g["L"]), float(g["lambda_"]))
    st.reset(g["prev0"][None])
    T = g["actions"].shape[0]
    for t in range(T + 1):
        pre = st.pretransition()
        np.testing.assert_array_equal(pre["obs"][0], g["obs"][t])
        np.testing.assert_array_equal(pre["beta"][0], g["beta"][t])
        if t < T:
            r, d = st.step(g["actions"][t][None])
            np.testing.assert_array_equal(r[0], g["rewards"][t])
            assert d == bool(g["done"][t])


def test_selectors_oracle_matches_reference():
    g = _load("selectors.npz")
    for t, e in zip(g["sched_t"], g["sched_eps"]):
        assert O.epsilon_linear(float(g["eps_start"]), float(g["eps_finish"]), float(g["eps_anneal"]), int(t)) == e
    eps_list = [O.epsilon_linear(1.0, 0.05, 1000, int(t)) for t in g["eg_t_env"]] + [float(g["eval_eps"])]
    for eps, want in zip(eps_list, g["eg_actions"]):
        got = O.select_epsilon_greedy(g["eg_q"], g["eg_avail"], eps, g["eg_u_explore"], g["eg_u_action"])
        np.testing.assert_array_equal(got, want)
    top = O.top_m_tasks(g["fg_beta"], int(g["fg_M"]))
    m = g["fg_beta"].shape[2]
    for eps, want in zip(eps_list, g["fg_actions"]):
        got = O.select_filtered_epsilon_greedy(g["fg_q"], top, g["eg_avail"], m, eps, g["fg_u_tie"],
                                               g["fg_u_explore"], g["fg_u_action"])
        np.testing.assert_array_equal(got, want)


def test_real_beta_hat_full_consistent():
    """beta_hat at the chosen entry equals the full tensor's entry (real_constellation_env.py:282-328)."""
    rng = np.random.default_rng(0)
    S = O.gen_ref_like(rng, 3, 6, 8, 5).astype(np.float64)
    beta = O.real_window(S, 1, 3)
    prev = rng.integers(0, 8, size=(3, 6))
    a = rng.integers(0, 8, size=(3, 6))
    full = O.real_beta_hat_full(beta, prev, 0.5)
    chosen = O.real_beta_hat_chosen(beta, prev, a, 0.5)
    np.testing.assert_array_equal(np.take_along_axis(full[..., 0], a[..., None], 2)[..., 0], chosen)


def test_rollout_timeline_real():
    """A.5 timeline: filled = 1 for t in [0,T], terminated only at T-1, final obs zeros."""
    rng = np.random.default_rng(5)
    S = O.gen_dense(rng, 2, 6, 8, 4).astype(np.float64)
    st = O.RealState(S, 3, 4, 2, 0.5)
    acts = rng.integers(0, 8, size=(4, 2, 6))
    out = O.rollout(st, lambda t, pre: acts[t], "real")
    assert out["filled"].sum() == 2 * 5
    assert out["terminated"][:, :3].sum() == 0 and out["terminated"][:, 3].all() and not out["terminated"][:, 4].any()
    assert not out["obs"][:, 4].any() and not out["beta"][:, 4].any()
    np.testing.assert_array_equal(out["prev_assigns"][:, 4], acts[3])


def test_sap_selectors_oracle_matches_reference():
    """The assignment selectors (sap_selectors.py, filtered_sap_selectors.py): the oracle's scipy restatement, fed the
    recorded Gaussian / tie draws, reproduces the reference's picks on every fixture."""
    g = _load("sap_selectors.npz")
    eps_list = [O.epsilon_linear(1.0, 0.05, 1000, int(t)) for t in g["sap_t_env"]] + [float(g["eval_eps"])]
    for eps, want in zip(eps_list, g["sap_actions"]):
        got, _ = O.lsa_maximize(g["sap_q"], g["sap_z"], O.sap_noise_std(g["sap_q"], eps))
        np.testing.assert_array_equal(got, want)
    got, _ = O.lsa_maximize(g["sap_q"])
    np.testing.assert_array_equal(got, g["egsap_test_actions"])
    M, m = int(g["sap_M"]), int(g["sap_m"])
    top = O.top_m_tasks(g["fsap_beta"], M)
    mat = O.filtered_benefit_matrix(g["fsap_q"], top, m, g["fsap_u_tie"])
    for eps, want in zip(eps_list, g["fsap_actions"]):
        got, _ = O.lsa_maximize(mat, g["fsap_z"], O.sap_noise_std(mat, eps))
        np.testing.assert_array_equal(got, want)
    got, _ = O.lsa_maximize(mat)
    np.testing.assert_array_equal(got, g["fepsgr_test_actions"])


def test_haa_oracle_matches_reference():
    """HAASelector of the reference (non_rl_selectors.py:10-50) along a short real-env episode."""
    g = _load("haa.npz")
    S = g["S"][None].astype(np.float64)
    st = O.RealState(S, int(g["L"]), int(g["M"]), int(g["N"]), float(g["lambda_"]))
    st.reset()
    for t, want in enumerate(g["haa_actions"]):
        got = O.haa_actions(st.beta, st.prev, float(g["lambda_"]))[0]
        np.testing.assert_array_equal(got, want)
        st.step((want if t % 2 == 0 else g["follow"][t])[None])


def test_policy_selectors_oracle_matches_reference():
    """Multinomial / SoftPolicies / FilteredSoftPolicies (classic_selectors.py:5-27, 56-64, filtered_classic_selectors.py:
    65-102) with the injected-uniform sampling contract."""
    g = _load("policy_selectors.npz")
    np.testing.assert_array_equal(O.sample_categorical(g["p"], g["u"], g["avail"]), g["multinomial_train"])
    np.testing.assert_array_equal((g["p"] * g["avail"]).argmax(-1), g["multinomial_test"])
    np.testing.assert_array_equal(O.sample_categorical(g["p"], g["u"]), g["soft"])
    M = int(g["M"])
    top = O.top_m_tasks(g["beta"], M)
    picked = O.sample_categorical(g["pf"], g["uf"])
    masked = g["u_rand"].copy()
    np.put_along_axis(masked, top, -1.0, axis=2)
    choices = np.concatenate([top, masked.argmax(-1)[..., None]], axis=2)
    np.testing.assert_array_equal(np.take_along_axis(choices, picked[..., None], axis=2)[..., 0], g["filtered"])


def test_bids_as_actions_oracle_matches_reference():
    """bids_as_actions: scipy assignment of the bid matrix, then the ordinary step (real and mock env)."""
    g = _load("bids.npz")
    S = g["S"][None].astype(np.float64)
    lam = float(g["lambda_"])
    st = O.RealState(S, int(g["L"]), int(g["M"]), int(g["N"]), lam)
    st.reset()
    np.testing.assert_array_equal(st.obs[0], g["real_obs"][0])
    for t, bids in enumerate(g["bids"]):
        a, _ = O.lsa_maximize(bids[None])
        r, _ = st.step(a)
        np.testing.assert_array_equal(r[0], g["real_rewards"][t])
        np.testing.assert_array_equal(st.obs[0], g["real_obs"][t + 1])
    ms = O.MockState(S, int(g["L"]), lam)
    ms.reset(g["mock_prev0"][None])
    for t, bids in enumerate(g["bids"]):
        r, _ = ms.step(O.lsa_maximize(bids[None])[0])
        np.testing.assert_array_equal(r[0], g["mock_rewards"][t])


def test_oracle_batched_rollout_matches_reference_parallel_runner():
    """B = 4 envs: the oracle's vectorised rollout (and its epsilon-greedy selector, fed the recorded draws and the Q-values
    of the recorded agent) against the batch the reference's own ParallelRunner wrote (runner_parallel.npz)."""
    import torch as th

    g = _load("runner_parallel.npz")
    B, n, m, T = int(g["B"]), int(g["n"]), int(g["m"]), int(g["T"])
    S = np.broadcast_to(g["S"].astype(np.float64), (B, n, m, T)).copy()
    st = O.RealState(S, int(g["L"]), int(g["M"]), int(g["N"]), float(g["lambda_"]))
    W = {k[2:]: th.tensor(v) for k, v in g.items() if k.startswith("w_")}
    last = [np.zeros((B, n, m), dtype=np.float32)]

    def policy(t, pre):
        obs = th.tensor(pre["obs"], dtype=th.float16).float().reshape(B * n, -1)
        x = th.cat([obs, th.tensor(last[0]).reshape(B * n, -1), th.eye(n).unsqueeze(0).expand(B, -1, -1).reshape(B * n, -1)], 1)
        h = th.relu(th.nn.functional.linear(x, W["fc1.weight"], W["fc1.bias"]))
        h = th.relu(th.nn.functional.linear(h, W["rnn.weight"], W["rnn.bias"]))
        q = th.nn.functional.linear(h, W["fc2.weight"], W["fc2.bias"]).reshape(B, n, m).numpy()
        a = O.select_epsilon_greedy(q, np.ones((B, n, m), bool), float(g["eps"]), g["u_explore"][t], g["u_action"][t])
        last[0] = O.one_hot(a, m, np.float32)
        return a

    want = O.rollout(st, policy, "real")
    np.testing.assert_array_equal(want["actions"][:, :T], g["td_actions"][:, :T, :, 0])
    np.testing.assert_array_equal(want["obs"].astype(np.float16), g["td_obs"])
    np.testing.assert_array_equal(want["rewards"].astype(np.float16), g["td_rewards"])
    np.testing.assert_array_equal(want["beta"].astype(np.float16), g["td_beta"])
    np.testing.assert_array_equal(want["prev_assigns"], g["td_prev_assigns"])
    np.testing.assert_array_equal(want["filled"], g["td_filled"][..., 0])
    # the reference ParallelRunner's two quirks, as recorded
    assert not g["td_terminated"][0].any() and g["td_terminated"][1:, :T].all() and not g["td_terminated"][:, T].any()
    assert g["td_actions"][:, T].any()
    assert np.mean(want["rewards"].sum((1, 2))) == pytest.approx(float(g["return_mean"]), rel=1e-12)


@pytest.mark.parametrize("name", ["power_env.npz", "interference_env.npz"])
def test_power_and_interference_oracle_matches_reference(name):
    """RealPowerConstellationEnv / InterferenceConstellationEnv: obs with the N + 1 power tail,
    rewards, the float64 power trajectory (incl. the 5.55e-17 residue that is `> 0` but `< 1e-12`, and the -0.2 it turns
    into) and done flags, against the unmodified reference classes."""
    g = _load(name)
    S = g["S"].astype(np.float64)[None]
    L, M, N, lam = int(g["L"]), int(g["M"]), int(g["N"]), float(g["lambda_"])
    if name.startswith("power"):
        st = O.PowerState(S, L, M, N, lam, task_prios=g["task_prios"])
    else:
        st = O.InterferenceState(S, L, M, N, lam, g["neighbor_matrix"], g["sat_freq_bands"], task_prios=g["task_prios"])
    assert st.obs_size == int(g["obs_size"])
    acts = g["actions"]
    want = O.rollout(st, lambda t, pre: acts[t][None], "real", prev0=g["prev0"][None])
    np.testing.assert_array_equal(want["obs"][0], g["obs"])
    np.testing.assert_array_equal(want["beta"][0], g["beta"])
    np.testing.assert_array_equal(want["prev_assigns"][0], g["prev"])
    np.testing.assert_array_equal(want["power_states"][0], g["power"])          # bit-exact float64
    np.testing.assert_allclose(want["rewards"][0, :acts.shape[0]], g["rewards"], rtol=1e-15, atol=0)
    assert want["terminated"][0, :acts.shape[0]].tolist() == [bool(d) for d in g["done"]]
    assert (g["power"][:, 5] < 0).any() and np.any((g["power"] > 0) & (g["power"] < 1e-12))


def test_haal_oracle_matches_reference():
    """HAALSelector of the unmodified reference (look-ahead over time-interval sequences) vs the oracle, step by step."""
    g = _load("haal.npz")
    S = g["S"].astype(np.float64)[None]
    assert O.time_interval_sequences(3) == [((0, 0), (1, 1), (2, 2)), ((0, 0), (1, 2)), ((0, 1), (2, 2)), ((0, 2),)]
    for t in range(S.shape[-1]):
        a = O.haal_actions(S, t, g["prev"][t][None], int(g["L"]), float(g["lambda_"]), task_prios=g["task_prios"])[0]
        np.testing.assert_array_equal(a, g["haal_actions"][t])
