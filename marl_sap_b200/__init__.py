"""marl_sap_b200: B200-native rollout hot path of josh-holder/marl_sap.

Env step, observation build, masked epsilon-greedy selection and episode-buffer insertion run as
hand-written sm_100a CUDA kernels behind the reference's registries:

    marl_sap_b200.envs.REGISTRY, marl_sap_b200.action_selectors.REGISTRY,
    marl_sap_b200.runners.REGISTRY, marl_sap_b200.components.episode_buffer.{EpisodeBatch, ReplayBuffer}

The kernels live in ``libmarl_sap_b200.so`` (C ABI: include/marl_sap_b200.h), built by
``python -m marl_sap_b200._build``.  There is no CPU fallback.
"""
__version__ = "0.1.0"
