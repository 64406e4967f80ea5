"""world_size = 2 gloo tests of the multi-rank host logic (run on CPU): the per-run statistics all-reduce of the
runner, the env block partition and the flat gradient all-reduce."""
import os
import socket
from types import SimpleNamespace

import numpy as np
import pytest
import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from marl_sap_b200.runners.vec_runner import CudaVecRunner, _Moments
    from marl_sap_b200.utils import dist as sdist
    from marl_sap_b200.utils.logging import Logger

    r, w, _ = sdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    # --- runner statistics: each rank owns B envs; return_mean/std must be over all ranks' episodes
    B, T = 3, 5
    runner = object.__new__(CudaVecRunner)
    runner.args = SimpleNamespace(test_nepisode=2 * B, runner_log_interval=1)
    runner.logger = Logger()
    runner.batch_size, runner.T, runner.t_env = B, T, 0
    runner.train_returns, runner.test_returns = _Moments(), _Moments()
    runner.train_stats, runner.test_stats = {}, {}
    runner.log_train_stats_t = -100000
    runner.mac = SimpleNamespace(action_selector=SimpleNamespace(epsilon=0.25))
    rets = th.tensor([1.0, 2.0, 4.0], dtype=th.float64) * (rank + 1)
    runner.last_episode_returns = rets
    runner._finish_run(test_mode=False)
    all_rets = np.concatenate([np.array([1.0, 2.0, 4.0]) * (k + 1) for k in range(world)])
    st = runner.logger.stats
    assert st["return_mean"][-1][1] == pytest.approx(all_rets.mean())
    assert st["return_std"][-1][1] == pytest.approx(all_rets.std())
    assert st["ep_length_mean"][-1][1] == T
    assert runner.t_env == world * B * T and st["steps"][-1][1] == world * B * T
    runner._finish_run(test_mode=True)  # 2B test episodes over 2 ranks = one run
    assert st["test_return_mean"][-1][1] == pytest.approx(all_rets.mean())
    # --- partition
    parts = [sdist.env_partition(10, k, 3) for k in range(3)]
    assert parts == [(0, 4), (4, 3), (7, 3)]
    # --- flat gradient all-reduce + broadcast
    lin = th.nn.Linear(3, 2)
    sdist.broadcast_parameters(lin, src=0)
    ref = [p.detach().clone() for p in lin.parameters()]
    gathered = [th.zeros_like(ref[0]) for _ in range(world)]
    dist.all_gather(gathered, ref[0])
    assert all(th.equal(g, gathered[0]) for g in gathered)
    for p in lin.parameters():
        p.grad = th.full_like(p, float(rank + 1))
    sdist.all_reduce_gradients(lin.parameters())
    for p in lin.parameters():
        assert th.allclose(p.grad, th.full_like(p, (1 + world) / 2))
    dist.barrier()
    dist.destroy_process_group()
    out.put(rank)


def test_two_rank_gloo_statistics_and_gradients():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0, "a rank failed"
    assert sorted(out.get(timeout=5) for _ in range(2)) == [0, 1]
