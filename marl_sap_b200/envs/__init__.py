"""Environment registry with the reference's keys (/root/reference/src/envs/__init__.py:19-30).

``REGISTRY[name](**env_args)`` returns a single-env object with the reference API backed by the CUDA
kernels.  ``BATCHED[name]`` is the batched device env the runners use for thousands of envs per launch.
"""
from functools import partial

from .batched import (BatchedInterferenceConstellationEnv, BatchedMockConstellationEnv, BatchedRealConstellationEnv,
                      BatchedRealPowerConstellationEnv)
from .single import InterferenceConstellationEnv, MockConstellationEnv, RealConstellationEnv, RealPowerConstellationEnv


def env_fn(env, **kwargs):
    return env(**kwargs)


def _not_on_path(name, why):
    def ctor(**kwargs):
        raise NotImplementedError(f"env '{name}' is not part of the B200 rollout hot path: {why} (DESIGN.md, 'out of scope')")
    return ctor


REGISTRY = {}
REGISTRY["mock_constellation_env"] = partial(env_fn, env=MockConstellationEnv)
REGISTRY["real_constellation_env"] = partial(env_fn, env=RealConstellationEnv)
# keys the reference registers that are outside SURVEY.md section 8 (a)-(e)
REGISTRY["real_power_constellation_env"] = partial(env_fn, env=RealPowerConstellationEnv)
REGISTRY["interference_constellation_env"] = partial(env_fn, env=InterferenceConstellationEnv)
REGISTRY["dictator_env"] = _not_on_path("dictator_env", "3x3 toy env")
REGISTRY["benefit_obs_env"] = _not_on_path("benefit_obs_env", "stale in the reference (no scheme)")
REGISTRY["power_constellation_env"] = _not_on_path("power_constellation_env", "stale in the reference (no scheme)")

BATCHED = {
    "mock_constellation_env": BatchedMockConstellationEnv,
    "real_constellation_env": BatchedRealConstellationEnv,
    "real_power_constellation_env": BatchedRealPowerConstellationEnv,
    "interference_constellation_env": BatchedInterferenceConstellationEnv,
}
