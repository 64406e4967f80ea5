"""Drop-in contract: the reference's OWN training driver and learner on top of this repo's runner, MAC and ReplayBuffer.

``run.run_sequential`` (/root/reference/src/run.py:107-324) and ``QLearner`` (learners/q_learner.py:11-191) are imported
UNMODIFIED from the reference tree (``oracle/_ref`` on the GPU box, see oracle/make_ref.py; the test is skipped when no
tree is available) and only the three names INTEGRATION.md says to swap are swapped: the runner registry, the MAC
registry and ``ReplayBuffer``.  Configuration = the reference's own yaml files for BASELINE.json configs[0]
(config/default.yaml + envs/mock_constellation_env.yaml + algs/mock_constellation_iql.yaml: IQL, jumpstart_mac with the
HAA jump-start policy, EpisodeRunner, standardised rewards, double Q) at 10 agents x 10 tasks, T = 100."""
import importlib
import logging
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch as th

pytestmark = pytest.mark.gpu


def _merged_config(ref_src, env_yaml, alg_yaml):
    import yaml

    def load(*parts):
        with open(os.path.join(ref_src, "config", *parts)) as f:
            return yaml.load(f, Loader=yaml.FullLoader)

    def update(d, u):  # main.py:61-67
        for k, v in u.items():
            d[k] = update(d.get(k, {}), v) if isinstance(v, dict) else v
        return d

    cfg = load("default.yaml")
    update(cfg, load("envs", env_yaml))
    update(cfg, load("algs", alg_yaml))
    return cfg


@pytest.mark.parametrize("runner_name,batch_size_run", [("episode", 1), ("parallel", 4)])
def test_reference_run_sequential_and_qlearner_drive_this_repo(runner_name, batch_size_run):
    from oracle import ref_import

    if not ref_import.reference_available():
        pytest.skip("no reference tree (run `python oracle/make_ref.py` where /root/reference exists)")
    ref_import.install()
    ref_run = importlib.import_module("run")
    ref_logging = importlib.import_module("utils.logging")

    from marl_sap_b200.components.episode_buffer import ReplayBuffer
    from marl_sap_b200.controllers import REGISTRY as mac_REGISTRY
    from marl_sap_b200.runners import REGISTRY as r_REGISTRY

    cfg = _merged_config(ref_import.REFERENCE_SRC, "mock_constellation_env.yaml", "mock_constellation_iql.yaml")
    cfg["env_args"].update(n=10, m=10, T=100, seed=1)
    cfg.update(seed=1, use_cuda=True, use_mps=False, device="cuda", buffer_cpu_only=False, use_mps_action_selection=True,
               runner=runner_name, batch_size_run=batch_size_run, t_max=1200, batch_size=4, buffer_size=8, test_nepisode=2 * batch_size_run,
               test_interval=600, log_interval=400, runner_log_interval=400, learner_log_interval=200, save_model=False,
               wandb_run_name="dropin", unique_token="dropin")
    args = SimpleNamespace(**cfg)
    logger = ref_logging.Logger(logging.getLogger("dropin"))

    saved = (ref_run.r_REGISTRY, ref_run.mac_REGISTRY, ref_run.ReplayBuffer)
    made = {}

    def mac_factory(scheme, groups, a, _cls=mac_REGISTRY[args.mac]):
        made["mac"] = _cls(scheme, groups, a)
        made["w0"] = {k: v.detach().clone() for k, v in made["mac"].agent.state_dict().items()}
        return made["mac"]

    def runner_factory(args, logger, _cls=r_REGISTRY[runner_name]):
        made["runner"] = _cls(args=args, logger=logger)
        return made["runner"]

    try:
        ref_run.r_REGISTRY = {runner_name: runner_factory}
        ref_run.mac_REGISTRY = {args.mac: mac_factory}
        ref_run.ReplayBuffer = ReplayBuffer
        th.manual_seed(1)
        np.random.seed(1)
        ref_run.run_sequential(args, logger)
    finally:
        ref_run.r_REGISTRY, ref_run.mac_REGISTRY, ref_run.ReplayBuffer = saved
    runner, mac = made["runner"], made["mac"]
    assert type(mac).__module__.startswith("marl_sap_b200") and type(runner).__module__.startswith("marl_sap_b200")
    assert runner.t_env > args.t_max and runner.t_env % (100 * batch_size_run) == 0
    stats = logger.stats
    for key in ("loss", "grad_norm", "td_error_abs", "q_taken_mean", "target_mean", "avg_num_conflicts", "avg_beta",
                "return_mean", "test_return_mean", "epsilon", "ep_length_mean"):
        assert key in stats and len(stats[key]) >= 1, key
        assert all(np.isfinite(float(v)) for _, v in stats[key]), key
    assert stats["ep_length_mean"][-1][1] == 100
    # the reference learner really trained this repo's agent through this repo's buffer
    moved = sum(float((v.cpu() - made["w0"][k].cpu()).abs().sum()) for k, v in mac.agent.state_dict().items())
    assert moved > 0
