"""Policy-sampling action selectors (agent_output_type == "pi_logits": the inputs are action probabilities).

MultinomialActionSelector      <- /root/reference/src/action_selectors/classic_selectors.py:5-27             ("multinomial")
SoftPoliciesSelector           <- /root/reference/src/action_selectors/classic_selectors.py:56-64            ("soft_policies")
FilteredSoftPoliciesSelector   <- /root/reference/src/action_selectors/filtered_classic_selectors.py:65-102  ("filtered_const_soft_policies")

``Categorical(p).sample()`` is one kernel, ``sap_sample_categorical``: inverse CDF in float64 on a uniform per
(env, agent).  torch.multinomial's stream cannot be reproduced, so the uniforms come from ``torch.rand`` on the device or
from ``inject_draws(u_sample=..., u_rand=...)`` in parity tests.  Inputs must be CUDA tensors; there is no CPU path.
"""
from __future__ import annotations

import torch as th

from .. import _lib
from .sap_selectors import _FilteredBase
from .selectors import _KernelSelectorBase


def sample_categorical(probs, avail=None, u=None):
    """One sample per row of ``probs[..., A]`` (masked by ``avail``): first k with cdf[k] > u * cdf[-1]."""
    _lib.require_cuda(probs, "probs")
    p = probs.detach()
    if p.dtype != th.float32:
        p = p.float()
    p = p.contiguous()
    A = p.shape[-1]
    rows = p.numel() // A
    if u is None:
        u = th.rand(rows, device=p.device, dtype=th.float32)
    u = u.to(device=p.device, dtype=th.float32).contiguous()
    assert u.numel() == rows
    av = _KernelSelectorBase._avail_u8(avail)
    out = th.empty(p.shape[:-1], dtype=th.int64, device=p.device)
    _lib.check(_lib.load().sap_sample_categorical(p.data_ptr(), _lib.ptr(av), rows, A, u.data_ptr(), out.data_ptr(),
                                                  _lib.stream_ptr(p.device)), "sap_sample_categorical")
    return out


class MultinomialActionSelector(_KernelSelectorBase):
    def __init__(self, args):
        super().__init__(args)
        self.test_greedy = getattr(args, "test_greedy", True)

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, beta=None):
        self.epsilon = self.schedule.eval(t_env)
        inj = self._injected or {}
        self._injected = None
        if test_mode and self.test_greedy:  # :23-24: arg-max of the masked policies (unavailable -> 0.0), first index
            masked = agent_inputs.detach().float()
            av = self._avail_u8(avail_actions)
            if av is not None:
                masked = masked * av.view_as(masked).to(masked.dtype)
            return masked.argmax(dim=2)
        return sample_categorical(agent_inputs, avail_actions, inj.get("u_sample"))  # :25-26


class SoftPoliciesSelector(_KernelSelectorBase):
    def __init__(self, args):
        self.args = args
        self.envs = None
        self._injected = None

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, beta=None):
        inj = self._injected or {}
        self._injected = None
        return sample_categorical(agent_inputs, None, inj.get("u_sample"))  # :61-63 (no masking in the reference)


class FilteredSoftPoliciesSelector(_FilteredBase):
    """Sample one of the M top tasks or the "anything else" slot; the latter becomes a uniformly random non-top task."""

    def __init__(self, args):
        self.args = args
        self.envs = None
        self._injected = None

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, beta=None, top=None):
        _lib.require_cuda(agent_inputs, "agent_inputs")
        q = agent_inputs.detach().float().contiguous()
        inj = self._injected or {}
        self._injected = None
        B, n, _ = q.shape
        top = self._top(q, beta, top).long()
        m = self._m(avail_actions, beta)
        picked = sample_categorical(q, None, inj.get("u_sample"))  # :77-78
        u_rand = inj.get("u_rand")
        if u_rand is None:
            u_rand = th.rand(B, n, m, device=q.device)  # :91
        masked = u_rand.to(q.device, th.float32).scatter(2, top, -1.0)  # :92-94
        baseline_task = masked.argmax(dim=-1, keepdim=True)  # :95
        choices = th.cat((top, baseline_task), dim=2)  # :97
        return choices.gather(2, picked.unsqueeze(-1)).squeeze(-1)  # :99-100
