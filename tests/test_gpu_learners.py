"""Device learners against the reference's own learners on identical batches and weights.

The reference classes are imported unmodified (``oracle/_ref`` on the GPU box; skipped without a reference tree) and are
given THIS repo's MAC and episode batch, which the drop-in test shows they accept; both learners then take the same
gradient step and must end with the same loss and the same parameters."""
import copy
import importlib
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch as th

from oracle import cpu_oracle as O

pytestmark = pytest.mark.gpu


def _ref_learner(name):
    from oracle import ref_import

    if not ref_import.reference_available():
        pytest.skip("no reference tree (oracle/_ref)")
    ref_import.install()
    return getattr(importlib.import_module("learners." + name), {"q_learner": "QLearner", "filtered_q_learner": "FilteredQLearner"}[name])


class _Log:
    def __init__(self):
        self.stats = {}

    def log_stat(self, k, v, t):
        self.stats.setdefault(k, []).append((t, float(v)))


def _setup(filtered, use_rnn=False, standardise_returns=False):
    from test_gpu_runner import build, make_args

    rng = np.random.default_rng(9)
    B, n, m, T, L, M, N = 6, 12, 16, 8, 3, 4, 3
    S = O.gen_dense(rng, B, n, m, T)
    if filtered:   # top-M action space: needs the real env's [n, m, L] beta
        env_name = "real_constellation_env"
        env_args = dict(num_planes=1, num_sats_per_plane=n, m=m, T=T, N=N, M=M, L=L, lambda_=0.5, sat_prox_mat=S, graphs=1)
    else:          # the reference's own IQL config runs on the mock env ([n, m] beta; its avg_beta diagnostic raises on 5-D)
        env_name = "mock_constellation_env"
        env_args = dict(n=n, m=m, T=T, L=L, lambda_=0.5, sat_prox_mat=S)
    args = make_args(env_name, env_args, B, agent="flat_const_agent" if filtered else "rnn",
                     selector="filtered_const_epsilon_greedy" if filtered else "epsilon_greedy", use_rnn=use_rnn,
                     epsilon_start=0.5, epsilon_finish=0.5, epsilon_anneal_time=1)
    args.__dict__.update(lr=5e-4, gamma=0.99, grad_norm_clip=10, double_q=True, mixer=None, use_cuda=True, use_mps=False,
                         standardise_rewards=True, standardise_returns=standardise_returns, target_update_interval_or_tau=0.01,
                         learner_log_interval=10 ** 9 if filtered else 1, optim_alpha=0.99, optim_eps=1e-5)
    runner, mac, buffer, _ = build(args)
    kw = {} if filtered else {"prev0": np.stack([rng.permutation(m)[:n] for _ in range(B)])}
    with th.no_grad():
        for _ in range(2):
            buffer.insert_episode_batch(runner.run(test_mode=False, **kw))
    batch = buffer.gather(list(range(2 * B)))
    return args, mac, buffer, batch


@pytest.mark.parametrize("filtered,use_rnn,std_ret", [(False, False, False), (False, False, True), (True, False, False), (False, True, False)])
def test_learner_step_matches_reference_learner(filtered, use_rnn, std_ret):
    from marl_sap_b200.learners import REGISTRY

    name = "filtered_q_learner" if filtered else "q_learner"
    RefLearner = _ref_learner(name)
    args, mac, buffer, batch = _setup(filtered, use_rnn, std_ret)
    mac_a, mac_b = copy.deepcopy(mac), copy.deepcopy(mac)
    log_a, log_b = _Log(), _Log()
    ours = REGISTRY[name](mac_a, buffer.scheme, log_a, args)
    ref = RefLearner(mac_b, buffer.scheme, log_b, args)
    if filtered:
        ref.log_stats_t = 10 ** 9   # never logs (see below); ours logs once
    ours.cuda()
    ref.cuda()
    for step in range(3):
        ours.train(batch, 100 * (step + 1), step)
        ref.train(batch, 100 * (step + 1), step)
    for key in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean", "avg_num_conflicts", "avg_beta"):
        a, b = [v for _, v in log_a.stats[key]], [v for _, v in log_b.stats.get(key, [])]
        if filtered:   # (logged once, at the first step: the reference's avg_beta raises on the real env's 5-D beta,
            assert len(a) == 1 and np.isfinite(a[0])   #  q_learner.py:172-179, so its log interval is set out of reach)
            continue
        assert len(a) == len(b) == 3, key
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=1e-6, err_msg=key)
    for (ka, pa), (kb, pb) in zip(mac_a.agent.state_dict().items(), mac_b.agent.state_dict().items()):
        assert ka == kb
        assert th.allclose(pa, pb, rtol=1e-4, atol=2e-6), ka
    for pa, pb in zip(ours.target_mac.parameters(), ref.target_mac.parameters()):
        assert th.allclose(pa, pb, rtol=1e-4, atol=2e-6)
    # the step moved the parameters at all
    assert any(not th.equal(p, q) for p, q in zip(mac_a.agent.parameters(), mac.agent.parameters()))


def test_running_mean_std_matches_reference():
    from oracle import ref_import

    if not ref_import.reference_available():
        pytest.skip("no reference tree (oracle/_ref)")
    ref_import.install()
    RefRMS = importlib.import_module("components.standarize_stream").RunningMeanStd
    from marl_sap_b200.components.standardize_stream import RunningMeanStd

    a, b = RunningMeanStd(shape=(5,), device="cuda"), RefRMS(shape=(5,), device="cuda")
    g = th.Generator(device="cuda").manual_seed(0)
    for _ in range(4):
        x = th.randn(7, 9, 5, device="cuda", generator=g) * 3 + 1
        a.update(x)
        b.update(x)
        assert th.allclose(a.mean, b.mean, rtol=1e-6, atol=1e-7) and th.allclose(a.var, b.var, rtol=1e-5) and a.count == pytest.approx(b.count)


def _nccl_worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    th.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=th.device("cuda", rank))
    from marl_sap_b200.components.standardize_stream import RunningMeanStd
    from marl_sap_b200.utils.dist import all_reduce_gradients, broadcast_parameters

    th.manual_seed(100 + rank)   # different weights and data per rank
    net = th.nn.Sequential(th.nn.Linear(6, 5), th.nn.ReLU(), th.nn.Linear(5, 3)).cuda()
    broadcast_parameters(net, src=0)
    w0 = th.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()
    x = th.randn(8, 6, device="cuda")
    net(x).pow(2).mean().backward()
    local = th.cat([p.grad.reshape(-1) for p in net.parameters()]).cpu()
    all_reduce_gradients(list(net.parameters()))
    avg = th.cat([p.grad.reshape(-1) for p in net.parameters()]).cpu()
    rms = RunningMeanStd(shape=(6,), device="cuda")
    rms.update(x)
    out[rank] = (w0, local, avg, x.cpu(), rms.mean.cpu(), rms.var.cpu())
    dist.destroy_process_group()


def test_gradient_all_reduce_and_broadcast_over_nccl():
    """utils.dist.all_reduce_gradients / broadcast_parameters and the cross-rank RunningMeanStd on two GPUs over NCCL
    (the CPU suite covers the same functions over gloo)."""
    if th.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, 29533, out), nprocs=2, join=True)
    (w0a, la, avga, xa, ma, va), (w0b, lb, avgb, xb, mb, vb) = out[0], out[1]
    assert th.equal(w0a, w0b)                                   # rank 0's weights everywhere
    assert th.allclose(avga, (la + lb) / 2, rtol=1e-6, atol=1e-8) and th.equal(avga, avgb)
    both = th.cat([xa, xb])
    assert th.allclose(ma, mb) and th.allclose(ma, both.mean(0) * 16 / (16 + 1e-4), rtol=1e-5, atol=1e-6)
    assert th.allclose(va, vb)
