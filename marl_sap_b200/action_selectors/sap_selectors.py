"""Assignment-problem action selectors backed by the batched linear-sum-assignment kernel (``sap_lsa_maximize``).

SequentialAssignmentProblemSelector    <- /root/reference/src/action_selectors/sap_selectors.py:52-97      ("sap")
EpsilonGreedySAPTestActionSelector     <- /root/reference/src/action_selectors/sap_selectors.py:7-48       ("epsilon_greedy_sap_test")
FilteredSAPActionSelector              <- /root/reference/src/action_selectors/filtered_sap_selectors.py:7-66   ("filtered_const_sap")
FilteredEpsGrSAPTestActionSelector     <- /root/reference/src/action_selectors/filtered_sap_selectors.py:68-149 ("filtered_const_epsgr_sap_test")

Same constructors and ``select_action(agent_inputs, avail_actions, t_env, test_mode=False, beta=None)`` signature.  The
reference loops over the batch on the host and calls ``scipy.optimize.linear_sum_assignment`` per env; here the whole
batch is one kernel launch (one warp per env, shortest augmenting paths on float64 duals, csrc/sap_lsa.cu).  Results
are returned as int64 ``[B, n]`` on the input's device (the reference returns a float tensor of the same values).

Random draws: the Gaussian perturbation ``th.normal`` and the tie-breaking ``th.rand_like`` streams of the reference
cannot be reproduced; draws come from ``torch.randn`` / ``torch.rand`` on the device, or from ``inject_draws(z=...,
u_tie=...)`` for parity tests.  Inputs must be CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import numpy as np
import torch as th

from .. import _lib
from .selectors import EpsilonGreedyActionSelector, FilteredEpsilonGreedyActionSelector, _KernelSelectorBase


def lsa_maximize(benefit, z=None, std=None, want_objective=False):
    """col_ind of ``scipy.optimize.linear_sum_assignment(benefit[b] + z[b] * std[b], maximize=True)`` for every b."""
    _lib.require_cuda(benefit, "benefit")
    q = benefit.detach()
    if q.dtype != th.float32:
        q = q.float()
    q = q.contiguous()
    B, n, m = q.shape
    if z is not None:
        z = z.to(device=q.device, dtype=th.float32).contiguous()
        std = std.to(device=q.device, dtype=th.float32).contiguous()
        assert tuple(z.shape) == (B, n, m) and std.numel() == B
    out = th.empty(B, n, dtype=th.int64, device=q.device)
    obj = th.empty(B, dtype=th.float64, device=q.device) if want_objective else None
    _lib.check(_lib.load().sap_lsa_maximize(q.data_ptr(), _lib.ptr(z), _lib.ptr(std), B, n, m, out.data_ptr(), _lib.ptr(obj),
                                            _lib.stream_ptr(q.device)), "sap_lsa_maximize")
    return (out, obj) if want_objective else out


def _noise_std(benefit, eps):
    """stds = ones * mean|benefit[b]| * eps * 2 (sap_selectors.py:84-85), one value per env, fp32 like the reference."""
    return benefit.abs().mean(dim=(1, 2)) * eps * 2


def filtered_benefit_matrix(q, top, m, u_tie=None):
    """[B, n, m] benefit matrix of the "top-M tasks + anything-else baseline" action space
    (filtered_sap_selectors.py:43-57): every entry = baseline + U * 1e-8, the top-M tasks get their own Q-values."""
    B, n, Mp1 = q.shape
    base = q[:, :, -1:].expand(B, n, m)
    if u_tie is None:
        u_tie = th.rand(B, n, m, device=q.device, dtype=th.float32)
    mat = base + u_tie.to(q.device, th.float32) * 1e-8
    mat.scatter_(2, top.long(), q[:, :, :-1])
    return mat


class SequentialAssignmentProblemSelector(_KernelSelectorBase):
    """Noise-perturbed Q matrix -> one optimal assignment per env ("sap")."""

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, beta=None):
        eps = self._eps(t_env, test_mode)  # test_mode: evaluation_epsilon (:62-64)
        q = agent_inputs.detach().float()
        inj = self._injected or {}
        self._injected = None
        z = inj.get("z")
        if z is None:
            z = th.randn(q.shape, device=q.device, dtype=th.float32)
        return lsa_maximize(q, z, _noise_std(q, eps))


class EpsilonGreedySAPTestActionSelector(_KernelSelectorBase):
    """Epsilon-greedy while training, an optimal assignment of the raw Q-values when testing."""

    def __init__(self, args):
        super().__init__(args)
        self._greedy = EpsilonGreedyActionSelector(args)

    def bind_counters(self, episode_ctr, step_k):
        super().bind_counters(episode_ctr, step_k)
        self._greedy.bind_counters(episode_ctr, step_k)

    def inject_draws(self, **draws):
        self._greedy.inject_draws(**draws)

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, beta=None):
        self.epsilon = self.schedule.eval(t_env)
        if test_mode:
            return lsa_maximize(agent_inputs)  # :26-33
        if np.random.rand() < self.epsilon:
            # :35-36 returns ONE th.randperm(n) for the whole batch (only well-formed when the batch is a single env);
            # here every env gets its own random permutation of the first n tasks
            B, n = agent_inputs.shape[0], agent_inputs.shape[1]
            return th.rand(B, n, device=agent_inputs.device).argsort(dim=1)
        out = self._greedy.select_action(agent_inputs, avail_actions, t_env, test_mode=False)  # :37-48
        self.epsilon = self._greedy.epsilon
        return out


class _FilteredBase(_KernelSelectorBase):
    def _top(self, q, beta, top):
        assert beta is not None or top is not None, "Need beta to figure out which are the top M tasks for each agent."
        B, n, Mp1 = q.shape
        M = self.args.env_args["M"]
        assert Mp1 == M + 1, f"agent must output M+1={M + 1} values, got {Mp1}"
        if top is not None:
            return top.contiguous()
        _lib.require_cuda(beta, "beta")
        bt = beta.contiguous()
        if bt.dtype not in (th.float32, th.float16):
            bt = bt.float()
        m, L = bt.shape[2], (bt.shape[3] if bt.dim() == 4 else 1)
        out = th.empty(B, n, M, dtype=th.int32, device=q.device)
        _lib.check(_lib.load().sap_topm_from_beta(bt.data_ptr(), _lib.sap_dtype(bt.dtype), B, n, m, L, M, out.data_ptr(),
                                                  _lib.stream_ptr(q.device)), "sap_topm_from_beta")
        return out

    def _m(self, avail_actions, beta):
        if beta is not None:
            return beta.shape[2]
        return avail_actions.shape[2] if avail_actions is not None else self.args.env_args["m"]


class FilteredSAPActionSelector(_FilteredBase):
    """"filtered_const_sap": the filtered benefit matrix, Gaussian perturbation, one optimal assignment per env."""

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, beta=None, top=None):
        eps = self._eps(t_env, test_mode)
        _lib.require_cuda(agent_inputs, "agent_inputs")
        q = agent_inputs.detach().float().contiguous()
        inj = self._injected or {}
        self._injected = None
        mat = filtered_benefit_matrix(q, self._top(q, beta, top), self._m(avail_actions, beta), inj.get("u_tie"))
        z = inj.get("z")
        if z is None:
            z = th.randn(mat.shape, device=mat.device, dtype=th.float32)
        return lsa_maximize(mat, z, _noise_std(mat, eps))


class FilteredEpsGrSAPTestActionSelector(_FilteredBase):
    """"filtered_const_epsgr_sap_test": filtered epsilon-greedy while training, an optimal assignment when testing."""

    def __init__(self, args):
        super().__init__(args)
        self._greedy = FilteredEpsilonGreedyActionSelector(args)

    def bind_counters(self, episode_ctr, step_k):
        super().bind_counters(episode_ctr, step_k)
        self._greedy.bind_counters(episode_ctr, step_k)

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, beta=None, top=None):
        self.epsilon = self.schedule.eval(t_env)
        if not test_mode:  # :112-149 (epsilon is NOT replaced by evaluation_epsilon here: test_mode never reaches it)
            if self._injected:
                self._greedy.inject_draws(**self._injected)
                self._injected = None
            return self._greedy.select_action(agent_inputs, avail_actions, t_env, test_mode=False, beta=beta, top=top)
        _lib.require_cuda(agent_inputs, "agent_inputs")
        q = agent_inputs.detach().float().contiguous()
        inj = self._injected or {}
        self._injected = None
        mat = filtered_benefit_matrix(q, self._top(q, beta, top), self._m(avail_actions, beta), inj.get("u_tie"))
        return lsa_maximize(mat)  # :84-111
