"""End-to-end rollout through the reference-facing API (runners.REGISTRY -> BasicMAC -> selector -> env kernels
-> EpisodeBatch -> ReplayBuffer) against the CPU oracle's EpisodeRunner-order rollout."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch as th

from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def make_args(env, env_args, B, selector="epsilon_greedy", agent="rnn", **kw):
    base = dict(env=env, env_args=env_args, batch_size_run=B, device="cuda", runner="parallel", mac="basic_mac",
                action_selector=selector, epsilon_start=1.0, epsilon_finish=0.05, epsilon_anneal_time=50000,
                evaluation_epsilon=0.0, agent=agent, hidden_dim=64, use_rnn=False, obs_agent_id=False,
                obs_last_action=False, agent_output_type="q", test_nepisode=B, runner_log_interval=1, seed=3,
                use_mps_action_selection=True)
    base.update(kw)
    return SimpleNamespace(**base)


def build(args):
    from marl_sap_b200.components.episode_buffer import ReplayBuffer
    from marl_sap_b200.controllers import REGISTRY as mac_REGISTRY
    from marl_sap_b200.runners import REGISTRY as r_REGISTRY
    from marl_sap_b200.utils.logging import Logger

    logger = Logger()
    runner = r_REGISTRY[args.runner](args=args, logger=logger)
    env = runner.get_env()
    args.n, args.m, args.T = env.n, env.m, env.T  # run.py:113-116
    groups = {"agents": args.n}
    buffer = ReplayBuffer(env.scheme, groups, 2 * args.batch_size_run + 1, env.T + 1, preprocess=env.preprocess,
                          device="cuda")
    th.manual_seed(0)
    mac = mac_REGISTRY[args.mac](buffer.scheme, groups, args)
    mac.cuda()
    runner.setup(scheme=env.scheme, groups=groups, preprocess=env.preprocess, mac=mac)
    return runner, mac, buffer, logger


class DrawInjector:
    """Feeds pre-generated uniforms to the selector so the oracle can replay the same decisions."""

    def __init__(self, selector, draws):
        self.sel, self.draws, self.t = selector, draws, 0
        self.orig = selector.select_action
        selector.select_action = self
        if hasattr(selector, "fused_select_args"):  # the selection inside the env step launch (sap_rollout_step)
            fused = selector.fused_select_args

            def hooked(*a, **k):
                self._inject()
                return fused(*a, **k)

            selector.fused_select_args = hooked

    def _inject(self):
        d = {name: th.tensor(v[self.t]).cuda() for name, v in self.draws.items()}
        self.sel.inject_draws(**d)
        self.t += 1

    def __call__(self, *a, **k):
        self._inject()
        return self.orig(*a, **k)


@pytest.mark.parametrize("env_name", ["real_constellation_env", "mock_constellation_env"])
@pytest.mark.parametrize("obs_extra", [False, True])
def test_runner_rollout_matches_oracle(env_name, obs_extra):
    rng = np.random.default_rng(17)
    B, n, m, T, L, M, N = 4, 12, 16, 9, 3, 4, 3
    S = O.gen_dense(rng, B, n, m, T)
    if env_name == "real_constellation_env":
        env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=0.5, sat_prox_mat=S, graphs=1)
        state = O.RealState(S.astype(np.float64), L, M, N, 0.5)
        kind, real_dt, int_dt = "real", th.float16, th.int16
    else:
        env_args = dict(n=n, m=m, T=T, L=L, lambda_=0.5, sat_prox_mat=S)
        state = O.MockState(S.astype(np.float64), L, 0.5)
        kind, real_dt, int_dt = "mock", th.float32, th.int64
    args = make_args(env_name, env_args, B, obs_agent_id=obs_extra, obs_last_action=obs_extra)
    runner, mac, buffer, logger = build(args)
    draws = {"u_explore": rng.random((T, B, n), dtype=np.float32), "u_action": rng.random((T, B, n), dtype=np.float32)}
    DrawInjector(mac.action_selector, draws)
    runner.t_env = 20000  # epsilon = 0.62: a mix of greedy and random picks
    eps = O.epsilon_linear(1.0, 0.05, 50000, 20000)
    prev0 = np.stack([rng.permutation(m)[:n] for _ in range(B)])

    last_onehot = [np.zeros((B, n, m), dtype=np.float32)]

    def policy(t, pre):
        obs = th.tensor(pre["obs"], dtype=real_dt).float()  # buffer rounding, then .float() (basic_controller.py:82)
        parts = [obs.reshape(B * n, -1)]
        if obs_extra:
            parts.append(th.tensor(last_onehot[0]).reshape(B * n, -1))
            parts.append(th.eye(n).unsqueeze(0).expand(B, -1, -1).reshape(B * n, -1))
        x = th.cat(parts, dim=1).cuda()
        with th.no_grad():
            q, _ = mac.agent(x, mac.agent.init_hidden().expand(B * n, -1))
        q = q.view(B, n, -1).cpu().numpy()
        a = O.select_epsilon_greedy(q, np.ones((B, n, m), bool), eps, draws["u_explore"][t], draws["u_action"][t])
        last_onehot[0] = O.one_hot(a, m, np.float32)
        return a

    want = O.rollout(state, policy, kind, prev0=prev0)
    with th.no_grad():
        batch = runner.run(test_mode=False, **({"prev0": prev0} if kind == "mock" else {}))
    td = {k: v.cpu() for k, v in batch.data.transition_data.items()}
    assert th.equal(td["actions"][..., 0], th.tensor(want["actions"]).to(int_dt))
    assert th.equal(td["obs"], th.tensor(want["obs"], dtype=real_dt))
    assert th.equal(td["rewards"], th.tensor(want["rewards"], dtype=real_dt))
    assert th.equal(td["beta"], th.tensor(want["beta"], dtype=real_dt))
    assert th.equal(td["terminated"][..., 0], th.tensor(want["terminated"]))
    assert th.equal(td["filled"][..., 0], th.tensor(want["filled"]))
    assert int(batch.max_t_filled()) == T + 1
    np.testing.assert_allclose(runner.last_episode_returns.cpu().numpy(), want["rewards"].sum((1, 2)), rtol=1e-12)
    # statistics contract (A.6): t_env += B*T, log keys
    assert runner.t_env == 20000 + B * T
    rets = want["rewards"].sum((1, 2))
    assert logger.stats["return_mean"][-1][1] == pytest.approx(np.mean(rets), rel=1e-9)
    assert logger.stats["return_std"][-1][1] == pytest.approx(np.std(rets), rel=1e-6)
    assert logger.stats["ep_length_mean"][-1][1] == T
    assert logger.stats["epsilon"][-1][1] == eps
    # ReplayBuffer insertion of the returned batch (run.py:262), twice to wrap the ring
    buffer.insert_episode_batch(batch)
    buffer.insert_episode_batch(batch)
    buffer.insert_episode_batch(batch)
    assert buffer.episodes_in_buffer == 2 * B + 1 and buffer.buffer_index == (3 * B) % (2 * B + 1)
    assert th.equal(buffer["obs"][B:2 * B].cpu(), td["obs"])
    assert buffer.can_sample(B)


def test_runner_filtered_selector_and_test_mode():
    """filtered_const_epsilon_greedy + flat_const_agent (filtered_iql.yaml) in greedy test mode."""
    rng = np.random.default_rng(23)
    B, n, m, T, L, M, N = 3, 10, 20, 6, 3, 4, 3
    S = O.gen_dense(rng, B, n, m, T)
    env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=0.5, sat_prox_mat=S, graphs=1)
    args = make_args("real_constellation_env", env_args, B, selector="filtered_const_epsilon_greedy", agent="flat_const_agent")
    runner, mac, buffer, logger = build(args)
    ut = rng.random((T, B, n, m), dtype=np.float32)
    draws = {"u_tie": ut, "u_explore": rng.random((T, B, n), dtype=np.float32),
             "u_action": rng.random((T, B, n), dtype=np.float32)}
    DrawInjector(mac.action_selector, draws)
    state = O.RealState(S.astype(np.float64), L, M, N, 0.5)

    def policy(t, pre):
        x = th.tensor(pre["obs"], dtype=th.float16).float().reshape(B * n, -1).cuda()
        with th.no_grad():
            q, _ = mac.agent(x, mac.agent.init_hidden().expand(B * n, -1))
        q = q.view(B, n, -1).cpu().numpy()
        top = O.top_m_tasks(pre["beta"], M)
        return O.select_filtered_epsilon_greedy(q, top, np.ones((B, n, m), bool), m, 0.0, ut[t], draws["u_explore"][t],
                                                draws["u_action"][t])

    want = O.rollout(state, policy, "real")
    with th.no_grad():
        batch = runner.run(test_mode=True)
    assert th.equal(batch["actions"][..., 0].cpu(), th.tensor(want["actions"]).to(th.int16))
    assert th.equal(batch["obs"].cpu(), th.tensor(want["obs"], dtype=th.float16))
    assert runner.t_env == 0  # test episodes do not advance t_env
    assert "test_return_mean" in logger.stats


def test_runner_lazy_fields_and_philox_reproducibility():
    rng = np.random.default_rng(29)
    B, n, m, T = 8, 10, 12, 7
    S = O.gen_ref_like(rng, B, n, m, T)
    env_args = dict(n=n, m=m, T=T, L=3, lambda_=0.5, sat_prox_mat=S)
    outs = []
    for rep in range(2):
        args = make_args("mock_constellation_env", env_args, B, lazy_buffer_fields=("avail_actions", "actions_onehot"),
                         epsilon_start=0.5, epsilon_finish=0.5)
        runner, mac, buffer, logger = build(args)
        prev0 = np.tile(np.arange(n), (B, 1))
        with th.no_grad():
            b1 = runner.run(prev0=prev0)
            a1 = b1["actions"].clone()
            b2 = runner.run(prev0=prev0)
        assert "avail_actions" not in b1.data.transition_data
        assert not th.equal(a1, b2["actions"])  # a new episode draws new randoms
        outs.append((a1.cpu(), b2["actions"].cpu()))
        oh = b2["actions_onehot"]
        assert oh.shape == (B, T + 1, n, m) and th.equal(oh[:, :T].argmax(-1, keepdim=True), b2["actions"][:, :T])
    assert th.equal(outs[0][0], outs[1][0]) and th.equal(outs[0][1], outs[1][1])  # same seed -> same rollout


def test_product_path_has_no_oracle_import():
    """The shipped package must not reach into oracle/ (prompt section 3)."""
    import os
    import re

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "marl_sap_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f"{f} imports the oracle"


def test_rollout_in_place_into_replay_ring():
    """runner.attach_replay: episodes land in the ring rows directly and equal a copy-inserted rollout."""
    rng = np.random.default_rng(31)
    B, n, m, T = 4, 8, 12, 5
    S = O.gen_dense(rng, B, n, m, T)
    env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=3, M=4, L=3, lambda_=0.5, sat_prox_mat=S, graphs=1)
    res = []
    for in_place in (False, True):
        args = make_args("real_constellation_env", env_args, B, epsilon_start=0.4, epsilon_finish=0.4)
        runner, mac, buffer, logger = build(args)  # ring of 2B + 1 rows
        if in_place:
            runner.attach_replay(buffer)
        for ep in range(3):  # third episode would wrap -> falls back to a private batch + copy insert
            batch = runner.run()
            if in_place and ep < 2:
                assert batch["obs"].data_ptr() == buffer["obs"][ep * B:].data_ptr()
            buffer.insert_episode_batch(batch)
        assert buffer.episodes_in_buffer == 2 * B + 1 and buffer.buffer_index == (3 * B) % (2 * B + 1)
        res.append({k: v.clone() for k, v in buffer.data.transition_data.items()})
    for k in res[0]:
        assert th.equal(res[0][k], res[1][k]), k


@pytest.mark.parametrize("env_name", ["real_constellation_env", "mock_constellation_env"])
def test_cuda_graph_rollout_equals_eager(env_name):
    """args.use_cuda_graph: the captured T-step loop reproduces the eager rollout bit for bit, episode after episode,
    while epsilon keeps following its schedule (device scalar)."""
    rng = np.random.default_rng(37)
    B, n, m, T = 6, 10, 12, 6
    S = O.gen_dense(rng, B, n, m, T)
    if env_name == "real_constellation_env":
        env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=3, M=4, L=3, lambda_=0.5, sat_prox_mat=S, graphs=1)
        kw = {}
    else:
        env_args = dict(n=n, m=m, T=T, L=3, lambda_=0.5, sat_prox_mat=S)
        kw = {"prev0": np.tile(np.arange(n), (B, 1))}
    outs = []
    for use_graph in (False, True):
        args = make_args(env_name, env_args, B, epsilon_anneal_time=200, reuse_episode_batch=True, use_cuda_graph=use_graph)
        runner, mac, buffer, logger = build(args)
        eps_seen, acts = [], []
        for ep in range(5):
            batch = runner.run(**kw)
            eps_seen.append(mac.action_selector.epsilon)
            acts.append(batch["actions"].clone())
            acts.append(batch["obs"].clone())
            acts.append(batch["rewards"].clone())
        if use_graph:
            assert len(runner._graphs) == 1
        outs.append((eps_seen, acts))
    assert outs[0][0] == outs[1][0] and outs[0][0][0] > outs[0][0][-1]  # epsilon annealed identically
    for a, b in zip(outs[0][1], outs[1][1]):
        assert th.equal(a, b)


@pytest.mark.parametrize("name", ["runner_mock", "runner_real"])
def test_episode_runner_matches_reference_runner_golden(name):
    """REGISTRY["episode"] runner + BasicMAC vs the EpisodeBatch produced by the UNMODIFIED reference EpisodeRunner +
    BasicMAC + RNNAgent + epsilon_greedy (tests/golden/make_golden.py: golden_runner), same weights, same injected
    draws, obs_last_action and obs_agent_id on.  Every buffer field must be identical."""
    import os

    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")))
    n, m, T = int(g["n"]), int(g["m"]), int(g["T"])
    if name == "runner_mock":
        env_name = "mock_constellation_env"
        env_args = dict(n=n, m=m, T=T, L=3, lambda_=0.5, sat_prox_mat=g["S"])
        kw = {"prev0": g["prev0"][None]}
    else:
        env_name = "real_constellation_env"
        env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, L=3, lambda_=0.5, N=int(g["N"]), M=int(g["M"]),
                        sat_prox_mat=g["S"], graphs=1)
        kw = {}
    args = make_args(env_name, env_args, 1, runner="episode", epsilon_start=0.4, epsilon_finish=0.4, epsilon_anneal_time=1,
                     obs_agent_id=True, obs_last_action=True)
    runner, mac, buffer, logger = build(args)
    mac.agent.load_state_dict({k[2:]: th.tensor(v) for k, v in g.items() if k.startswith("w_")})
    DrawInjector(mac.action_selector, {"u_explore": g["u_explore"], "u_action": g["u_action"]})
    batch = runner.run(test_mode=False, **kw)
    for k, v in batch.data.transition_data.items():
        want = th.tensor(g["td_" + k])
        assert th.equal(v.cpu(), want), f"field {k} differs from the reference runner's batch"
    assert runner.t_env == int(g["t_env_after"])
    assert logger.stats["return_mean"][-1][1] == pytest.approx(float(g["return_mean"]), rel=1e-3)  # fp16 rewards in the log sum


@pytest.mark.parametrize("in_dim,hidden,n_out,rows", [(490, 64, 100, 4096), (301, 32, 11, 777), (40, 64, 10, 10)])
def test_agent_rollout_forward_matches_the_reference_module_math(in_dim, hidden, n_out, rows):
    """The rollout path of the agent (mm + sap_bias_act / fused-ReLU addmm, cached W^T) computes the reference's
    RNNAgent.forward (modules/agents/rnn_agent.py:22-31) in fp32: equal to the F.relu(F.linear(...)) chain up to the
    accumulation order cuBLAS picks per shape (1e-5 relative, the north star's fp32 tolerance)."""
    import torch.nn.functional as F

    from marl_sap_b200.modules.agents import RNNAgent

    th.manual_seed(in_dim)
    args = SimpleNamespace(hidden_dim=hidden, use_rnn=False, m=n_out)
    agent = RNNAgent(in_dim, args).cuda()
    x = th.randn(rows, in_dim, device="cuda")
    h0 = agent.init_hidden().expand(rows, -1)

    def ref():
        h1 = F.relu(F.linear(x, agent.fc1.weight, agent.fc1.bias))
        h2 = F.relu(F.linear(h1, agent.rnn.weight, agent.rnn.bias))
        return F.linear(h2, agent.fc2.weight, agent.fc2.bias), h2

    with th.no_grad():
        q, h = agent(x, h0)
        q_ref, h_ref = ref()
    th.testing.assert_close(h, h_ref, rtol=1e-5, atol=1e-6)
    th.testing.assert_close(q, q_ref, rtol=1e-5, atol=1e-6)
    # an in-place parameter update must invalidate the cached transposes
    with th.no_grad():
        agent.fc1.weight.mul_(0.5)
        q2, _ = agent(x, h0)
        q2_ref, _ = ref()
    th.testing.assert_close(q2, q2_ref, rtol=1e-5, atol=1e-6)
    assert not th.allclose(q2, q)


@pytest.mark.parametrize("rows,cols,relu", [(4096, 64, 1), (777, 33, 1), (5, 100, 0), (1, 4, 1)])
def test_bias_act_kernel_is_bit_exact(rows, cols, relu):
    """sap_bias_act (in-place x = act(x + bias)) against torch, bit for bit, vector and scalar paths."""
    from marl_sap_b200 import _lib

    g = th.Generator(device="cuda").manual_seed(rows * 131 + cols)
    x = th.randn(rows, cols, device="cuda", generator=g)
    b = th.randn(cols, device="cuda", generator=g)
    want = x + b
    if relu:
        want = th.relu(want)
    _lib.check(_lib.load().sap_bias_act(x.data_ptr(), b.data_ptr(), rows, cols, relu, _lib.stream_ptr()), "sap_bias_act")
    assert th.equal(x, want)


@pytest.mark.parametrize("selector,agent", [("sap", "rnn"), ("filtered_const_sap", "flat_const_agent")])
def test_runner_with_assignment_selectors_matches_oracle(selector, agent):
    """iql_sap.yaml / filtered_reda.yaml style rollouts: the runner with the assignment selectors against the oracle
    rollout (scipy per env) fed the same Gaussian / tie draws; every env's joint action is conflict-free."""
    rng = np.random.default_rng(41)
    B, n, m, T, L, M, N = 3, 10, 20, 5, 3, 4, 3
    S = O.gen_dense(rng, B, n, m, T)
    env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=0.5, sat_prox_mat=S, graphs=1)
    args = make_args("real_constellation_env", env_args, B, selector=selector, agent=agent, epsilon_start=0.3,
                     epsilon_finish=0.3, epsilon_anneal_time=1)
    runner, mac, buffer, logger = build(args)
    draws = {"z": rng.standard_normal((T, B, n, m)).astype(np.float32), "u_tie": rng.random((T, B, n, m), dtype=np.float32)}
    if selector == "sap":
        del draws["u_tie"]
    DrawInjector(mac.action_selector, draws)
    state = O.RealState(S.astype(np.float64), L, M, N, 0.5)

    def policy(t, pre):
        x = th.tensor(pre["obs"], dtype=th.float16).float().reshape(B * n, -1).cuda()
        with th.no_grad():
            q, _ = mac.agent(x, mac.agent.init_hidden().expand(B * n, -1))
        q = q.view(B, n, -1).cpu().numpy()
        if selector == "sap":
            mat = q
        else:
            mat = O.filtered_benefit_matrix(q, O.top_m_tasks(pre["beta"], M), m, draws["u_tie"][t])
        return O.lsa_maximize(mat, draws["z"][t], O.sap_noise_std(mat, 0.3))[0]

    want = O.rollout(state, policy, "real")
    with th.no_grad():
        batch = runner.run(test_mode=False)
    acts = batch["actions"][..., 0].cpu()
    assert th.equal(acts, th.tensor(want["actions"]).to(th.int16))
    assert th.equal(batch["rewards"].cpu(), th.tensor(want["rewards"], dtype=th.float16))
    for t in range(T):
        for b in range(B):
            assert len(set(acts[b, t].tolist())) == n  # an assignment: no two agents on one task


def test_haa_selector_and_jumpstart_mac():
    """haa_selector + jumpstart_mac (filtered_reda.yaml): HAA picks equal the reference's on its golden episode, through
    the env-bound path and through the batch's own state fields; a jump-started rollout matches the oracle."""
    import os

    from marl_sap_b200.action_selectors.non_rl_selectors import REGISTRY as non_rl

    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "haa.npz")))
    S = g["S"].astype(np.float32)
    n, m, T = S.shape
    L, M, N, lam = int(g["L"]), int(g["M"]), int(g["N"]), float(g["lambda_"])
    B = 3
    env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=lam, sat_prox_mat=S, graphs=1)
    args = make_args("real_constellation_env", env_args, B, selector="filtered_const_sap", agent="flat_const_agent",
                     mac="jumpstart_mac", jumpstart_action_selector="haa_selector", jumpstart_epsilon_start=1.0,
                     jumpstart_epsilon_finish=1.0, jumpstart_epsilon_anneal_time=1, jumpstart_evaluation_epsilon=1.0)
    runner, mac, buffer, logger = build(args)
    assert type(mac).__name__ == "JumpstartMAC"
    # 1. pure HAA rollout (jumpstart epsilon = 1): every step the optimal assignment of beta_hat
    st = O.RealState(np.broadcast_to(S, (B, n, m, T)).astype(np.float64), L, M, N, lam)
    want = O.rollout(st, lambda t, pre: O.haa_actions(pre["beta"], pre["prev_assigns"], lam), "real")
    with th.no_grad():
        batch = runner.run(test_mode=False)
    assert th.equal(batch["actions"][..., 0].cpu(), th.tensor(want["actions"]).to(th.int16))
    assert th.equal(batch["rewards"].cpu(), th.tensor(want["rewards"], dtype=th.float16))
    np.testing.assert_array_equal(want["actions"][0, 0], g["haa_actions"][0])  # the reference's own first pick
    # 2. the reference's golden episode step by step, once via the bound env and once via the batch fields
    for bound in (True, False):
        runner.reset()
        sel = non_rl["haa_selector"](args)
        if bound:
            sel.bind_env(runner.env)
        else:
            sel._env = None
            sel.args.env_args = env_args
        for t, want_a in enumerate(g["haa_actions"]):
            if bound:
                got = sel.select_action(runner.batch, t)
            else:  # generic path needs prev_assigns / beta in the batch (beta is materialised lazily)
                got = sel.select_action(runner.batch, t)
            np.testing.assert_array_equal(got[0].cpu().numpy(), want_a)
            a = want_a if t % 2 == 0 else g["follow"][t]
            runner.env.step(th.tensor(np.broadcast_to(a, (B, n)).copy(), device="cuda"), runner.batch)
    assert hasattr(non_rl["haal_selector"](args), "select_action")   # built: see the HAAL test below


def test_runner_mock_env_draws_its_own_benefits_on_the_device():
    """MockConstellationEnv without sat_prox_mat (the reference redraws its benefits at every reset,
    mock_constellation_env.py:99-100): the batched runner generates them on the device, new ones every episode."""
    args = make_args("mock_constellation_env", dict(n=6, m=8, T=12, L=3, lambda_=0.5), 5)
    runner, mac, buffer, logger = build(args)
    assert not runner.env.constant_benefits
    with th.no_grad():
        b1 = runner.run(test_mode=False)
        p1 = runner.env.planes.clone()
        o1 = b1["obs"].clone()
        b2 = runner.run(test_mode=False)
    assert not th.equal(p1, runner.env.planes) and not th.equal(o1, b2["obs"])
    assert th.equal(o1[:, 0, :, 8:16].cpu(), p1[:, 0].cpu())  # first obs slice = benefits at k = 0
    assert bool(th.isfinite(b2["rewards"].float()).all()) and int(b2["filled"].sum()) == 5 * 13


def test_rollout_diagnostics_match_the_learner_loops():
    """Device versions of q_learner.py:157-191 (conflicts per joint action, mean chosen benefit) vs the loops."""
    from marl_sap_b200.utils.rollout_stats import (calc_conflicting_actions, calc_raw_benefits,
                                                   calc_raw_benefits_from_planes)

    rng = np.random.default_rng(3)
    B, T, n, m = 4, 7, 9, 6
    acts = rng.integers(0, m, size=(B, T, n))
    beta = rng.random((B, T, n, m)).astype(np.float32)
    a = th.tensor(acts).cuda().unsqueeze(-1).to(th.int16)
    assert calc_conflicting_actions(a, m) == pytest.approx(O.calc_conflicting_actions(acts, m), rel=1e-12)
    want = O.calc_raw_benefits(beta.astype(np.float64), acts)
    assert calc_raw_benefits(th.tensor(beta).cuda(), a) == pytest.approx(want, rel=1e-9)
    assert calc_raw_benefits(th.tensor(beta).cuda().unsqueeze(-1).expand(-1, -1, -1, -1, 3), a) == pytest.approx(want, rel=1e-9)
    assert calc_raw_benefits_from_planes(th.tensor(beta).cuda(), a) == pytest.approx(want, rel=1e-9)


def test_agent_fc1_split_precision_matches_fp32():
    """Opt-in first layer (args.agent_fc1 = "fp16_split"): fp16 rows against [W_0 | 2^-11 W_1 | 2^-22 W_2] on the tensor
    cores + sap_split_bias_act == relu(F.linear(x.float(), W, b)) to accumulation-order accuracy (2e-6 relative)."""
    from marl_sap_b200.modules.agents import RNNAgent

    th.manual_seed(4)
    args = SimpleNamespace(hidden_dim=64, use_rnn=False, m=100, agent_fc1_terms=3)
    agent = RNNAgent(490, args).cuda()
    rows = 4096
    x16 = th.zeros(rows, 496, dtype=th.float16, device="cuda")
    x16[:, :490] = (th.rand(rows, 490, device="cuda") * 3).half()
    x16[:, :490][th.rand(rows, 490, device="cuda") < 0.3] = 0
    with th.no_grad():
        want = th.relu(th.nn.functional.linear(x16[:, :490].double(), agent.fc1.weight.double(), agent.fc1.bias.double()))
        got = agent._linear_fp16_split(agent.fc1, x16, relu=True)
        ref32 = th.relu(th.nn.functional.linear(x16[:, :490].float(), agent.fc1.weight, agent.fc1.bias))
        scale = want.abs().max().item()
        err_split = (got.double() - want).abs().max().item() / scale
        err_fp32 = (ref32.double() - want).abs().max().item() / scale
        assert err_split < 2e-6, (err_split, err_fp32)
        # two pieces: 22 mantissa bits of W
        agent.args.agent_fc1_terms = 2
        got2 = agent._linear_fp16_split(agent.fc1, x16, relu=True)
        assert (got2.double() - want).abs().max().item() / scale < 2e-5
        # whole forward: fp16 rows in, same Q as the fp32 path
        agent.args.agent_fc1_terms = 3
        q16, _ = agent(x16, agent.init_hidden().expand(rows, -1))
        q32, _ = agent(x16[:, :490].float(), agent.init_hidden().expand(rows, -1))
        assert (q16 - q32).abs().max().item() < 5e-6 * max(1.0, q32.abs().max().item())
    # an in-place weight update invalidates the cached pieces
    with th.no_grad():
        agent.fc1.weight.mul_(0.5)
        got3 = agent._linear_fp16_split(agent.fc1, x16, relu=False)
        want3 = th.nn.functional.linear(x16[:, :490].double(), agent.fc1.weight.double(), agent.fc1.bias.double())
        assert (got3.double() - want3).abs().max().item() / want3.abs().max().item() < 2e-6


@pytest.mark.parametrize("extra", [False, True])
def test_runner_fp16_split_agent_matches_default(extra):
    """The opt-in rollout (fp16 staging rows, split-precision fc1) against the default one (fp32 staging rows, sgemm) at
    the bench shape: identical observations / rewards as long as the actions agree, and Q-values equal to 1e-5."""
    rng = np.random.default_rng(23)
    B, n, m, T, L, M, N = 6, 100, 100, 4, 3, 10, 10
    S = O.gen_dense(rng, B, n, m, T)
    env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=0.5, sat_prox_mat=S, graphs=1)
    draws = {"u_explore": rng.random((T, B, n), dtype=np.float32), "u_action": rng.random((T, B, n), dtype=np.float32)}
    out = {}
    for mode in ("fp32", "fp16_split"):
        args = make_args("real_constellation_env", env_args, B, agent_fc1=mode, obs_agent_id=extra, obs_last_action=extra,
                         epsilon_start=0.3, epsilon_finish=0.3)
        runner, mac, buffer, _ = build(args)
        assert runner.agent_in.dtype == (th.float32 if mode == "fp32" else th.float16)
        assert runner.agent_in.shape[-1] == ((490 + (200 if extra else 0) + 7) // 8 * 8 if mode == "fp16_split" else 490 + (200 if extra else 0))
        DrawInjector(mac.action_selector, draws)
        qs = []
        orig = mac.forward

        def fwd(*a, _o=orig, _q=qs, **k):
            q = _o(*a, **k)
            _q.append(q.clone())
            return q

        mac.forward = fwd
        with th.no_grad():
            batch = runner.run(test_mode=False)
        out[mode] = (batch["obs"].clone(), batch["actions"].clone(), batch["rewards"].clone(), th.stack(qs))
    (o32, a32, r32, q32), (o16, a16, r16, q16) = out["fp32"], out["fp16_split"]
    same = (a32 == a16).all(dim=2).all(dim=2)  # [B, T+1]: envs whose joint action agrees at every step so far
    assert same[:, 0].all() or (q32[0] - q16[0]).abs().max() < 1e-5
    assert (q32[0] - q16[0]).abs().max().item() < 1e-5 * max(1.0, q32[0].abs().max().item())
    agree = same.cumprod(dim=1).bool()
    assert agree[:, 0].float().mean() > 0.8  # near-ties may flip a greedy pick; the bulk must agree
    for b in range(B):
        for t in range(T):
            if agree[b, :t + 1].all():
                assert th.equal(o32[b, t + 1], o16[b, t + 1]) and th.equal(r32[b, t], r16[b, t])


def test_runner_graph_refuses_jumpstart_and_tracks_buffers():
    """CUDA-graph rollout: a MAC that decides on the host (JumpstartMAC) is never captured; a BasicMAC capture is keyed on
    every buffer address and keeps its buffers alive; test-mode episodes never roll out over replay rows."""
    rng = np.random.default_rng(5)
    B, n, m, T, L = 3, 6, 8, 5, 2
    S = O.gen_ref_like(rng, B, n, m, T)
    env_args = dict(n=n, m=m, T=T, L=L, lambda_=0.5, sat_prox_mat=S)
    args = make_args("mock_constellation_env", env_args, B, use_cuda_graph=True)
    runner, mac, buffer, _ = build(args)
    runner.attach_replay(buffer)
    prev0 = np.stack([rng.permutation(m)[:n] for _ in range(B)])
    with th.no_grad():
        b1 = runner.run(test_mode=False, prev0=prev0)
        buffer.insert_episode_batch(b1)
        b2 = runner.run(test_mode=False, prev0=prev0)   # second episode: captured
        buffer.insert_episode_batch(b2)
        assert len(runner._graphs) == 1
        (key, entry), = runner._graphs.items()
        assert entry[1] is b2 and len(key) >= len(b2.data.transition_data) + 3
        rows_before = {k: v.clone() for k, v in buffer.data.transition_data.items()}
        bt = runner.run(test_mode=True, prev0=prev0)    # test episode: private batch, ring untouched
        assert getattr(bt, "_ring_owner", None) is None
        for k, v in buffer.data.transition_data.items():
            assert th.equal(v, rows_before[k]), k
    # JumpstartMAC: refused
    from marl_sap_b200.controllers import REGISTRY as mac_REGISTRY

    args2 = make_args("mock_constellation_env", env_args, B, use_cuda_graph=True, mac="jumpstart_mac",
                      jumpstart_action_selector="haa_selector", jumpstart_epsilon_start=0.5, jumpstart_epsilon_finish=0.5,
                      jumpstart_epsilon_anneal_time=1, jumpstart_evaluation_epsilon=0.0)
    runner2, mac2, _, _ = build(args2)
    assert isinstance(mac2, mac_REGISTRY["jumpstart_mac"]) and not mac2.graph_capturable
    with th.no_grad():
        for _ in range(3):
            runner2.run(test_mode=False, prev0=prev0)
    assert not getattr(runner2, "_graphs", {})


@pytest.mark.parametrize("compat", [True, False])
def test_parallel_runner_matches_reference_parallel_runner_golden(compat):
    """REGISTRY["parallel"] with B = 4 envs against the batch the reference's own ParallelRunner produced
    (tests/golden/runner_parallel.npz: forked env workers, injected draws).  Every field the reference's bugs do not touch
    is bit-identical; with ``compat_parallel_runner_quirks`` the two quirky ones are too (`terminated` from the list
    truthiness, the extra selection stored at t = T); without it they follow EpisodeRunner semantics."""
    import os

    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "runner_parallel.npz")))
    B, n, m, T = int(g["B"]), int(g["n"]), int(g["m"]), int(g["T"])
    env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=int(g["N"]), M=int(g["M"]), L=int(g["L"]),
                    lambda_=float(g["lambda_"]), sat_prox_mat=g["S"].astype(np.float64), graphs=1, seed=0)
    args = make_args("real_constellation_env", env_args, B, obs_agent_id=True, obs_last_action=True, epsilon_start=0.4,
                     epsilon_finish=0.4, epsilon_anneal_time=1, compat_parallel_runner_quirks=compat)
    runner, mac, buffer, logger = build(args)
    mac.agent.load_state_dict({k[2:]: th.tensor(v) for k, v in g.items() if k.startswith("w_")})
    DrawInjector(mac.action_selector, {"u_explore": g["u_explore"], "u_action": g["u_action"]})
    env = runner.get_env()   # a12: the single-env facade with the reference env API, not just a shape record
    assert hasattr(env, "beta_hat") and hasattr(env, "step") and (env.n, env.m, env.T) == (n, m, T)
    with th.no_grad():
        batch = runner.run(test_mode=False)
    td = {k: v.cpu() for k, v in batch.data.transition_data.items()}
    for k in ("obs", "rewards", "beta", "prev_assigns", "avail_actions", "filled"):
        assert th.equal(td[k], th.tensor(g["td_" + k])), k
    assert th.equal(td["actions"][:, :T], th.tensor(g["td_actions"][:, :T]))
    assert th.equal(td["actions_onehot"][:, :T], th.tensor(g["td_actions_onehot"][:, :T]))
    if compat:
        assert th.equal(td["actions"], th.tensor(g["td_actions"]))                 # incl. the selection stored at t = T
        assert th.equal(td["actions_onehot"], th.tensor(g["td_actions_onehot"]))
        assert th.equal(td["terminated"], th.tensor(g["td_terminated"]))           # env 0: never, others: always
    else:
        assert not td["actions"][:, T].any() and not td["actions_onehot"][:, T].any()
        want = th.zeros(B, T + 1, 1, dtype=th.bool)
        want[:, T - 1] = True
        assert th.equal(td["terminated"], want)
    assert runner.t_env == int(g["t_env_after"])
    assert logger.stats["return_mean"][-1][1] == pytest.approx(float(g["return_mean"]), rel=1e-9)


@pytest.mark.parametrize("env_name", ["real_constellation_env", "mock_constellation_env"])
def test_lazy_replay_buffer_rebuilds_beta_for_every_stored_episode(env_name):
    """A ReplayBuffer that does not store `beta` (lazy) still returns the right `beta` for every episode it holds - from
    `buffer["beta"]`, from `sample` / `gather` and from time slices - also after the env's benefits were replaced between
    episodes, and it drops the planes of a generation once the ring has overwritten every episode that used them."""
    from marl_sap_b200.components.episode_buffer import ReplayBuffer

    rng = np.random.default_rng(31)
    B, n, m, T, L, M, N = 3, 12, 16, 6, 3, 4, 3
    real = env_name == "real_constellation_env"
    S = [O.gen_dense(rng, B, n, m, T) for _ in range(3)]
    if real:
        env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=0.5, sat_prox_mat=S[0], graphs=1)
    else:
        env_args = dict(n=n, m=m, T=T, L=L, lambda_=0.5, sat_prox_mat=S[0])
    lazy = ("beta", "avail_actions", "actions_onehot")
    args = make_args(env_name, env_args, B, lazy_buffer_fields=lazy)
    runner, mac, _, _ = build(args)
    env = runner.get_env()
    buf = ReplayBuffer(env.scheme, {"agents": n}, 2 * B, T + 1, preprocess=env.preprocess, device="cuda", lazy=lazy)
    assert "beta" not in buf.data.transition_data
    runner.attach_replay(buf)
    dt = th.float16 if real else th.float32

    def want_beta(Sx):
        st = O.RealState(Sx.astype(np.float64), L, M, N, 0.5) if real else O.MockState(Sx.astype(np.float64), L, 0.5)
        acts = rng.integers(0, m, size=(T, B, n))
        kw = {} if real else {"prev0": np.stack([rng.permutation(m)[:n] for _ in range(B)])}
        return th.tensor(O.rollout(st, lambda t, pre: acts[t], "real" if real else "mock", **kw)["beta"], dtype=dt)

    wants = [want_beta(s) for s in S]
    kw = {} if real else {"prev0": np.stack([rng.permutation(m)[:n] for _ in range(B)])}
    with th.no_grad():
        buf.insert_episode_batch(runner.run(test_mode=False, **kw))
        runner.env.load_benefits(S[1])                      # the env's benefits change between the episodes
        ep2 = runner.run(test_mode=False, **kw)
        assert th.equal(ep2["beta"].cpu(), wants[1])        # the runner's own batch
        buf.insert_episode_batch(ep2)
    both = th.cat([wants[0], wants[1]])
    assert th.equal(buf["beta"].cpu(), both)
    assert th.equal(buf.sample(2 * B)["beta"].cpu(), both)
    ids = [4, 1, 3]
    picked = buf.gather(ids)
    assert th.equal(picked["beta"].cpu(), both[ids])
    assert th.equal(picked[:, 2:5]["beta"].cpu(), both[ids][:, 2:5])
    assert th.equal(buf[1:5, 1:]["beta"].cpu(), both[1:5, 1:])
    assert bool(picked["avail_actions"].all()) and picked["actions_onehot"].shape == (3, T + 1, n, m)
    src = buf.benefit_source
    assert len(src.generations) == 2
    with th.no_grad():
        runner.env.load_benefits(S[2])
        buf.insert_episode_batch(runner.run(test_mode=False, **kw))   # wraps: overwrites the rows of the first episode
    assert th.equal(buf["beta"].cpu(), th.cat([wants[2], wants[1]]))
    assert len(src.generations) == 2 and 0 not in src.generations   # generation 0 is no longer referenced: dropped


def test_replay_insert_of_a_stale_ring_view_does_not_alias():
    """insert_episode_batch of a view_next() batch taken BEFORE the ring moved on: source and destination rows alias the
    same storage, so the insert goes through a private copy instead of an overlapping row copy."""
    from marl_sap_b200.components.episode_buffer import ReplayBuffer

    scheme = {"obs": {"vshape": 5, "group": "agents", "dtype": th.float32}, "rewards": {"vshape": (1,), "dtype": th.float32}}
    buf = ReplayBuffer(scheme, {"agents": 2}, 6, 4, device="cuda")
    v = buf.view_next(3)                      # rows 0..2
    v.data.transition_data["obs"].copy_(th.arange(3 * 4 * 2 * 5, dtype=th.float32, device="cuda").view(3, 4, 2, 5))
    other = ReplayBuffer(scheme, {"agents": 2}, 2, 4, device="cuda")
    buf.insert_episode_batch(other[0:2])      # the ring moves on by two rows (host path: views)
    assert buf.buffer_index == 2
    keep = v["obs"].clone()
    buf.insert_episode_batch(v)               # stale view: lands in rows 2..4, overlapping its own storage (rows 0..2)
    assert th.equal(buf["obs"][2:5], keep) and buf.buffer_index == 5


def test_haal_selector_matches_reference_golden_and_oracle():
    """haal_selector on the batched env: the reference's picks along its golden episode (tests/golden/haal.npz) and the
    oracle's on a batch of distinct envs."""
    import os

    from marl_sap_b200.action_selectors.non_rl_selectors import REGISTRY as non_rl
    from marl_sap_b200.components.episode_buffer import EpisodeBatch
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv

    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "haal.npz")))
    S = g["S"]
    n, m, T = S.shape
    L, M, N, lam = int(g["L"]), int(g["M"]), int(g["N"]), float(g["lambda_"])
    env = BatchedRealConstellationEnv(1, n, m, T, L, M, N, lam, sat_prox_mat=S, task_prios=g["task_prios"])
    batch = EpisodeBatch(env.scheme, {"agents": n}, 1, T + 1, preprocess=env.preprocess, device="cuda")
    sel = non_rl["haal_selector"](SimpleNamespace(runner="episode"))
    sel.bind_env(env)
    env.reset(batch)
    for t in range(T):
        a = sel.select_action(batch, t)
        assert a.cpu().numpy()[0].tolist() == g["haal_actions"][t].tolist(), t
        env.step(a if t % 2 == 0 else th.tensor(g["follow"][t][None], device="cuda"), batch)
    # a batch of distinct envs against the oracle
    rng = np.random.default_rng(3)
    B, n, m, T = 5, 10, 14, 6
    S = O.gen_dense(rng, B, n, m, T)
    env = BatchedRealConstellationEnv(B, n, m, T, 3, 4, 3, 0.5, sat_prox_mat=S)
    batch = EpisodeBatch(env.scheme, {"agents": n}, B, T + 1, preprocess=env.preprocess, device="cuda")
    sel.bind_env(env)
    env.reset(batch)
    for t in range(T):
        want = O.haal_actions(S, t, env.prev.cpu().numpy(), 3, 0.5)
        a = sel.select_action(batch, t)
        np.testing.assert_array_equal(a.cpu().numpy(), want)
        env.step(a, batch)


@pytest.mark.parametrize("mode", ["eager", "graph", "fp16_split", "filtered"])
def test_runner_overlapped_obs_build_equals_fused_step(mode):
    """args.overlap_obs_build (default): the observation of slot t + 1 is built on a second stream next to the agent forward
    of step t (sap_real_obs_ahead) and the step only adds rewards / counters / flags (sap_real_step_after_obs).  Every
    field of the episode batch must equal the fused-step rollout bit for bit - eagerly, as a captured CUDA graph, with the
    fp16 staging rows of the split-precision agent, and with the filtered selector (which reads the env's top-M tasks)."""
    rng = np.random.default_rng(29)
    B, n, m, T, L, M, N = 5, 100, 100, 5, 3, 10, 10
    S = O.gen_dense(rng, B, n, m, T)
    S[:, :, ::9] = 0.0
    env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=0.5, sat_prox_mat=S, graphs=1)
    draws = {"u_explore": rng.random((2 * T, B, n), dtype=np.float32), "u_action": rng.random((2 * T, B, n), dtype=np.float32)}
    if mode == "filtered":
        draws["u_tie"] = rng.random((2 * T, B, n, m), dtype=np.float32)
    out = []
    for overlap in (False, True):
        kw = dict(overlap_obs_build=overlap, epsilon_start=0.3, epsilon_finish=0.3, obs_last_action=True, obs_agent_id=True,
                  use_cuda_graph=mode == "graph", reuse_episode_batch=True)
        if mode == "fp16_split":
            kw["agent_fc1"] = "fp16_split"
        if mode == "filtered":
            kw.update(agent="flat_const_agent", selector="filtered_const_epsilon_greedy")
        args = make_args("real_constellation_env", env_args, B, **kw)
        runner, mac, buffer, _ = build(args)
        assert runner._overlap == overlap
        inj = None if mode == "graph" else DrawInjector(mac.action_selector, draws)   # graphs: in-kernel Philox (same seeds)
        eps = []
        with th.no_grad():
            for _ in range(3 if mode == "graph" else 2):
                b = runner.run(test_mode=False)
                eps.append({k: v.clone() for k, v in b.data.transition_data.items()})
                if inj is not None and len(eps) == 1:
                    inj.t = T   # second episode: the second half of the draws
        out.append((eps, runner.env.ep_return.clone(), runner.env.top.clone(), runner.env.prev.clone()))
        if mode == "graph":
            assert len(runner._graphs) == 1
    (ea, ra, ta, pa), (eb, rb, tb, pb) = out
    for x, y in zip(ea, eb):
        for k in x:
            assert th.equal(x[k], y[k]), k
    assert th.equal(ra, rb) and th.equal(ta, tb) and th.equal(pa, pb)
    assert eb[0]["obs"][:, 1:, :, -M:].any() and not th.equal(eb[0]["obs"], eb[1]["obs"])


@pytest.mark.parametrize("mode", ["injected", "philox", "graph"])
@pytest.mark.parametrize("overlap", [False, True])
def test_runner_fused_select_step_equals_two_launches(mode, overlap):
    """args.fuse_select_step (default): the epsilon-greedy selection runs inside the env step launch (sap_rollout_step,
    SURVEY.md 7.4).  Every buffer field, the env state and the returns equal the selector kernel + step kernel schedule,
    with injected draws and with the in-kernel Philox stream, eagerly and replayed from a CUDA graph."""
    rng = np.random.default_rng(77)
    B, n, m, T, L, M, N = 5, 100, 100, 5, 3, 10, 10
    S = O.gen_dense(rng, B, n, m, T)
    env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=0.5, sat_prox_mat=S, graphs=1)
    draws = {"u_explore": rng.random((2 * T, B, n), dtype=np.float32), "u_action": rng.random((2 * T, B, n), dtype=np.float32)}
    out = []
    for fuse in (False, True):
        args = make_args("real_constellation_env", env_args, B, overlap_obs_build=overlap,
                         fuse_select_step=("always" if fuse else False),
                         epsilon_start=0.4, epsilon_finish=0.4, obs_agent_id=True, use_cuda_graph=mode == "graph",
                         reuse_episode_batch=True)
        runner, mac, buffer, _ = build(args)
        inj = DrawInjector(mac.action_selector, draws) if mode == "injected" else None
        eps = []
        with th.no_grad():
            for _ in range(3 if mode == "graph" else 2):
                b = runner.run(test_mode=False)
                eps.append({k: v.clone() for k, v in b.data.transition_data.items()})
                if inj is not None and len(eps) == 1:
                    inj.t = T
        assert runner._fused_select() == fuse
        runner.args.fuse_select_step = True   # the default fuses only in front of the light step of the overlapped schedule
        assert runner._fused_select() == overlap
        out.append((eps, runner.env.ep_return.clone(), runner.env.top.clone(), runner.env.prev.clone(), runner.env.counts.clone()))
    (ea, *sa), (eb, *sb) = out
    for x, y in zip(ea, eb):
        for k in x:
            assert th.equal(x[k], y[k]), k
    for x, y in zip(sa, sb):
        assert th.equal(x, y)
    acts = eb[0]["actions"][:, :T, :, 0].long()
    assert acts.min() >= 0 and acts.max() < m and len(th.unique(acts)) > m // 2
    assert not th.equal(eb[0]["actions"], eb[1]["actions"])


def test_rollout_step_entry_point_matches_selector_plus_step_and_the_oracle():
    """sap_rollout_step through the C ABI: actions == oracle epsilon-greedy on the same Q and draws, and the step it performs
    == sap_real_step on those actions (both schedules: full step, and after sap_real_obs_ahead)."""
    from marl_sap_b200 import _lib
    from marl_sap_b200.components.episode_buffer import EpisodeBatch
    from marl_sap_b200.envs.batched import BatchedRealConstellationEnv, real_obs_size, real_scheme

    rng = np.random.default_rng(5)
    B, n, m, T, L, M, N = 3, 100, 100, 4, 3, 10, 10
    S = O.gen_dense(rng, B, n, m, T)
    scheme, preprocess = real_scheme(n, m, L, real_obs_size(M, N, L))
    eps = 0.35

    def fresh():
        env = BatchedRealConstellationEnv(B, n, m, T, L, M, N, 0.5, sat_prox_mat=S, device="cuda")
        batch = EpisodeBatch(scheme, {"agents": n}, B, T + 1, preprocess=preprocess, device="cuda")
        env.reset(batch)
        return env, batch

    for ahead in (False, True):
        (env_a, batch_a), (env_b, batch_b) = fresh(), fresh()
        state = O.RealState(S.astype(np.float64), L, M, N, 0.5)
        state.reset()
        for t in range(T):
            q = rng.standard_normal((B, n, m)).astype(np.float32)
            q[:, :, 7] = q[:, :, 3]  # ties: first index wins
            ue, ua = rng.random((B, n), dtype=np.float32), rng.random((B, n), dtype=np.float32)
            want = O.select_epsilon_greedy(q, np.ones((B, n, m), bool), eps, ue, ua)
            qd, ued, uad = th.tensor(q).cuda(), th.tensor(ue).cuda(), th.tensor(ua).cuda()
            sel = _lib.SapSelectArgs()
            sel.q, sel.u_explore, sel.u_action, sel.eps, sel.seed = qd.data_ptr(), ued.data_ptr(), uad.data_ptr(), eps, 1
            acts = th.full((B, n), -1, dtype=th.int64, device="cuda")
            if ahead:
                env_a.obs_ahead(batch_a)
                env_b.obs_ahead(batch_b)
            env_a.step_select(sel, acts, batch_a)
            np.testing.assert_array_equal(acts.cpu().numpy(), want)
            env_b.step(th.tensor(want).cuda(), batch_b)
            state.step(want)
        for k in batch_a.data.transition_data:
            assert th.equal(batch_a.data.transition_data[k], batch_b.data.transition_data[k]), k
        assert th.equal(env_a.ep_return, env_b.ep_return) and th.equal(env_a.prev, env_b.prev) and th.equal(env_a.k, env_b.k)
        np.testing.assert_array_equal(env_a.prev.cpu().numpy(), state.prev)
