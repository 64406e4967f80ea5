"""Import shims that make the UNMODIFIED reference importable in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``marl_sap_b200``) may import
this module; it is used by ``tests/golden/make_golden.py`` (fixture generation, run in the
container that has ``/root/reference``) and by CPU tests that are skipped when the
reference tree is absent (it does not exist on the GPU box).

What is stubbed:
  * ``gym``  -- only the names touched at import/ctor time by
    /root/reference/src/envs/__init__.py:12-17 and mock_constellation_env.py:5-7,56-63.
    Stub classes are module-level so that the env objects stay picklable
    (parallel_runner.py:281-282 pickles the env over a Pipe).
  * ``envs.HighPerformanceConstellationSim`` -- only used when no ``sat_prox_mat`` is
    passed (real_constellation_env.py:47-53); needs poliastro/h3 which are absent.
  * ``matplotlib``, ``astropy`` -- module-level imports of utils/methods.py:3,5.
  * ``sacred``/``wandb``/``tensorboard_logger`` are never imported on the path we use.
"""
from __future__ import annotations

import os
import sys
import types

def _reference_root() -> str:
    """`/root/reference` in the build container; on the GPU box the copy `oracle/make_ref.py` left under `oracle/_ref`."""
    cands = [os.environ.get("MARL_SAP_REFERENCE"), "/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "src", "envs")):
            return c
    return cands[1]


REFERENCE_ROOT = _reference_root()
REFERENCE_SRC = os.path.join(REFERENCE_ROOT, "src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "envs"))


# --------------------------------------------------------------------------- gym stub
class _Space:
    def __init__(self, *a, **k):
        self.args = a
        self.kwargs = k
        self.shape = k.get("shape", None)


class Box(_Space):
    pass


class Discrete(_Space):
    def __init__(self, n):
        super().__init__(n)
        self.n = n


class Tuple(_Space):  # noqa: A001 - mirrors gym.spaces.Tuple
    def __init__(self, spaces):
        super().__init__()
        self.spaces = tuple(spaces)

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)


def flatdim(space):
    return 0


def flatten(space, x):
    return x


class Env:
    pass


class Wrapper(Env):
    def __init__(self, env=None):
        self.env = env


class ObservationWrapper(Wrapper):
    pass


class TimeLimit(Wrapper):
    pass


def _np_random(seed=None):
    import numpy as np

    return np.random.RandomState(seed), seed


def _make(*a, **k):
    raise RuntimeError("gym stub: gym.make is not available")


def _register(*a, **k):
    return None


def _install_gym_stub():
    if "gym" in sys.modules and not getattr(sys.modules["gym"], "_marl_sap_stub", False):
        return  # a real gym is present; use it
    me = sys.modules[__name__]
    gym = types.ModuleType("gym")
    gym._marl_sap_stub = True
    gym.Env = Env
    gym.Wrapper = Wrapper
    gym.ObservationWrapper = ObservationWrapper
    gym.make = _make
    gym.register = _register

    spaces = types.ModuleType("gym.spaces")
    for name in ("Box", "Discrete", "Tuple", "flatdim", "flatten"):
        setattr(spaces, name, getattr(me, name))
    gym.spaces = spaces

    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = _np_random
    utils.seeding = seeding
    gym.utils = utils

    envs = types.ModuleType("gym.envs")
    envs.registry = {}
    envs.registration = types.ModuleType("gym.envs.registration")
    envs.registration.register = _register
    gym.envs = envs

    wrappers = types.ModuleType("gym.wrappers")
    wrappers.TimeLimit = TimeLimit
    gym.wrappers = wrappers

    sys.modules.update({
        "gym": gym, "gym.spaces": spaces, "gym.utils": utils, "gym.utils.seeding": seeding,
        "gym.envs": envs, "gym.envs.registration": envs.registration, "gym.wrappers": wrappers,
    })


class HighPerformanceConstellationSim:  # stub: orbit propagation is out of scope
    def __init__(self, *a, **k):
        raise RuntimeError("HighPerformanceConstellationSim stub: pass sat_prox_mat and graphs")


def _install_misc_stubs():
    for name in ("matplotlib", "matplotlib.pyplot", "astropy", "astropy.units"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])
    if "astropy" in sys.modules and "astropy.units" in sys.modules:
        setattr(sys.modules["astropy"], "units", sys.modules["astropy.units"])


_installed = False


def install():
    """Put the reference's ``src`` on sys.path behind the stubs. Idempotent."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_SRC}")
    _install_gym_stub()
    _install_misc_stubs()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    # The reference's `envs` package imports .HighPerformanceConstellationSim (poliastro, h3).
    hp = types.ModuleType("envs.HighPerformanceConstellationSim")
    hp.HighPerformanceConstellationSim = HighPerformanceConstellationSim
    sys.modules["envs.HighPerformanceConstellationSim"] = hp
    _installed = True


def ref_modules():
    """Return the reference modules on the hot path (imports them on first use)."""
    install()
    import importlib

    names = {
        "envs": "envs",
        "real_env": "envs.real_constellation_env",
        "mock_env": "envs.mock_constellation_env",
        "classic_selectors": "action_selectors.classic_selectors",
        "filtered_selectors": "action_selectors.filtered_classic_selectors",
        "sap_selectors": "action_selectors.sap_selectors",
        "filtered_sap_selectors": "action_selectors.filtered_sap_selectors",
        "selectors": "action_selectors",
        "episode_buffer": "components.episode_buffer",
        "transforms": "components.transforms",
        "schedules": "components.epsilon_schedules",
        "runners": "runners",
        "episode_runner": "runners.episode_runner",
        "parallel_runner": "runners.parallel_runner",
        "basic_controller": "controllers.basic_controller",
        "agents": "modules.agents",
    }
    return types.SimpleNamespace(**{k: importlib.import_module(v) for k, v in names.items()})


class stable_argsort:
    """Context manager: force ``numpy.argsort`` default kind to "stable" inside the reference.

    The reference calls ``np.argsort(x)`` with numpy's unstable default at
    real_constellation_env.py:198,206,217, whose tie order depends on the CPU's SIMD sort.
    Parity on tie-heavy inputs is defined against the stable rule.
    """

    def __enter__(self):
        import numpy as np

        self._np = np
        self._orig = np.argsort

        def argsort(a, axis=-1, kind=None, order=None, **kw):
            return self._orig(a, axis=axis, kind="stable" if kind is None else kind, order=order, **kw)

        np.argsort = argsort
        return self

    def __exit__(self, *exc):
        self._np.argsort = self._orig
        return False
