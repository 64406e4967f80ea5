"""Benefit tensors from constellation geometry, on the device.

``fov_proximities`` = /root/reference/src/envs/HighPerformanceConstellationSim.py ``calc_fov_based_proximities_fast``
(:308-327) over every (satellite, task, time step), i.e. the ``sat_prox_mat`` that ``get_proximities_for_random_tasks``
(:91-173) and ``get_proximities_for_coverage_tasks`` (:175-270) build with Python triple loops.  The orbit propagation that
yields the satellite positions (poliastro) is out of scope: positions are an input, ``sat_r[n, 3, T]`` in km like the
simulator's ``sat_rs_over_time``.  The result lands directly in the env kernels' plane layout, so
``env.set_planes(fov_proximities(...)[None], shared=True)`` needs no re-layout and no host round trip."""
from __future__ import annotations

import math

import numpy as np
import torch as th

from .. import _lib


def gaussian_sigma_2(fov=60.0, prox_at_max_fov=0.05):
    """:100-101 - the Gaussian falls to ``prox_at_max_fov`` at the edge of the field of view."""
    return math.sqrt(-(fov ** 2) / (2 * math.log(prox_at_max_fov))) ** 2


def fov_proximities(sat_r, task_r, fov=60.0, prox_at_max_fov=0.05, device="cuda", reference_layout=False):
    """planes [T, n, m] fp32 on the device (and, with ``reference_layout``, also sat_prox_mat [n, m, T] float64)."""
    sat = th.as_tensor(np.asarray(sat_r) if not isinstance(sat_r, th.Tensor) else sat_r, dtype=th.float64).contiguous().to(device)
    task = th.as_tensor(np.asarray(task_r) if not isinstance(task_r, th.Tensor) else task_r, dtype=th.float64).contiguous().to(device)
    if sat.dim() != 3 or sat.shape[1] != 3 or task.dim() != 2 or task.shape[1] != 3:
        raise ValueError(f"sat_r must be [n, 3, T] and task_r [m, 3], got {tuple(sat.shape)} and {tuple(task.shape)}")
    n, _, T = sat.shape
    m = task.shape[0]
    planes = th.empty(T, n, m, dtype=th.float32, device=sat.device)
    ref = th.empty(n, m, T, dtype=th.float64, device=sat.device) if reference_layout else None
    _lib.check(_lib.load().sap_proximities_fov(sat.data_ptr(), task.data_ptr(), n, m, T, float(fov),
                                               gaussian_sigma_2(fov, prox_at_max_fov), planes.data_ptr(), _lib.ptr(ref),
                                               _lib.stream_ptr(sat.device)), "sap_proximities_fov")
    return (planes, ref) if reference_layout else planes
