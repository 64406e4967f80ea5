"""Device versions of the learner-side rollout diagnostics (first piece).

/root/reference/src/learners/q_learner.py:157-191 computes them with Python triple loops over (batch, time, agent) and
``.item()`` on every action; here they are a handful of device array ops on the episode batch's own tensors.

calc_conflicting_actions(actions, m)  <- q_learner.py:157-170  mean over (b, t) of sum_j max(count_j - 1, 0)
calc_raw_benefits(beta, actions)      <- q_learner.py:172-191  mean over (b, t, i) of beta[b, t, i, a_i] (4-D beta; the
                                         reference's 5-D branch tests ``beta.dim == 5`` - a method, never true - and
                                         raises; here 5-D beta means [..., L] and the current step l = 0 is used, which
                                         is what that branch says it wants)
calc_raw_benefits_from_planes(planes, actions)  same quantity straight from the env's planes [B, T, n, m] (no beta field)
"""
from __future__ import annotations

import torch as th


def calc_conflicting_actions(actions, m=None):
    a = actions.long()
    if a.dim() == 4:
        a = a[..., 0]
    B, T, n = a.shape
    srt = a.sort(dim=2).values
    distinct = 1 + (srt[..., 1:] != srt[..., :-1]).sum(-1)  # tasks chosen at least once
    return float((n - distinct).sum().item()) / B / T      # duplicates = n - distinct


def calc_raw_benefits(beta, actions):
    if beta.dim() == 5:
        beta = beta[..., 0]
    elif beta.dim() != 4:
        raise ValueError("beta has unexpected shape.")
    a = actions.long()
    if a.dim() == 3:
        a = a.unsqueeze(-1)
    B, T, n = a.shape[:3]
    chosen = beta[:, :T].gather(3, a)
    return float(chosen.double().sum().item()) / B / T / n


def calc_raw_benefits_from_planes(planes, actions, task_prios=None):
    """planes [B or 1, T, n, m] fp32 (the env's benefit layout); actions [B, T', n(, 1)] with T' <= T."""
    a = actions.long()
    if a.dim() == 3:
        a = a.unsqueeze(-1)
    B, T, n = a.shape[:3]
    p = planes[:, :T]
    if p.shape[0] == 1 and B > 1:
        p = p.expand(B, -1, -1, -1)
    chosen = p.gather(3, a).double()
    if task_prios is not None:
        chosen = chosen * task_prios.double()[a.squeeze(-1)].unsqueeze(-1)
    return float(chosen.sum().item()) / B / T / n
