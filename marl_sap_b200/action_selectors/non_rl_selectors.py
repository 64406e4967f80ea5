"""Non-learning jump-start policies (/root/reference/src/action_selectors/non_rl_selectors.py).

HAASelector ("haa_selector", :10-50): every env takes the optimal assignment of its CURRENT handover-aware benefits,
``linear_sum_assignment(beta_hat(beta, prev_assigns)[..., 0], maximize=True)``.  The reference loops over the batch on
the host (``env.beta_hat`` + scipy per env); here the whole batch is a few device array ops for beta_hat (the formula of
real_constellation_env.py:282-328 / mock_constellation_env.py:228-274) and one launch of the batched assignment kernel.
Like the reference it works on the ``beta`` / ``prev_assigns`` state fields of the episode batch, i.e. on benefits
already rounded to the scheme dtype; when the runner has bound its batched env (``bind_env``) the same values are
taken from the env's planes instead of materialising the lazy ``beta`` field.

HAALSelector ("haal_selector", :54-145) searches over time-interval sequences with deep-copied envs; not built
(DESIGN.md section 9).
"""
from __future__ import annotations

import torch as th

from .sap_selectors import lsa_maximize


def beta_hat_now(beta, prev, lambda_, T_trans=None):
    """beta_hat[..., 0] (real env, beta [B,n,m,L]) or beta_hat (mock env, beta [B,n,m]) in float64 on the device."""
    b = beta.double()
    B, n, m = b.shape[:3]
    prev = prev.long().reshape(B, n)
    if T_trans is None:
        pen = (th.arange(m, device=b.device).view(1, 1, m) != prev.unsqueeze(-1)).double()  # default T_trans = 1 - I
    else:
        pen = T_trans.double()[prev]  # onehot(prev) @ T_trans (:304-314)
    if b.dim() == 4:
        meaningful = (b.sum(-1) > 1e-12).double()  # :317
        return b[..., 0] - lambda_ * pen * meaningful
    return b - lambda_ * pen * (b > 1e-12).double()  # mock env: per-element test (:263-270)


class HAASelector:
    def __init__(self, args):
        self.args = args
        self.envs = None   # assigned by the runners, like the learning selectors' (episode_runner.py:39)
        self._env = None   # the batched device env, when bound

    def bind_env(self, env):
        self._env = env

    def _state_now(self, batch, t):
        env = self._env
        if env is not None and env.planes is not None and env.t_host == t:
            # the window the env kernel would expose at this step, rounded to the buffer dtype exactly like batch["beta"]
            dt = batch.scheme["beta"]["dtype"]
            k = env.t_host
            if env.kind == "real":
                L = env.L
                win = env.planes[:, k:k + L]
                if win.shape[1] < L:
                    pad = th.zeros(win.shape[0], L - win.shape[1], env.n, env.m, device=win.device, dtype=win.dtype)
                    win = th.cat([win, pad], dim=1)
                beta = win.permute(0, 2, 3, 1).double()
                if env.task_prios is not None:
                    beta = beta * env.task_prios.double().view(1, 1, -1, 1)
            else:
                beta = (env.planes[:, k] if k < env.T else th.zeros_like(env.planes[:, 0])).double()
            beta = beta.to(dt)
            if beta.shape[0] == 1 and env.B > 1:
                beta = beta.expand(env.B, *beta.shape[1:])
            return beta, env.prev, env.lambda_, env.T_trans
        # generic path: the state fields of the episode batch, as in the reference (:35-39)
        lam = env.lambda_ if env is not None else self.args.env_args["lambda_"]
        T_trans = env.T_trans if env is not None else None
        if "prev_assigns" in batch.scheme and env is None:
            prev = batch["prev_assigns"][:, t]
        else:
            prev = env.prev  # the mock env never writes prev_assigns into the batch (mock_constellation_env.py:164-175)
        return batch["beta"][:, t], prev, lam, T_trans

    def select_action(self, batch, t=0):
        beta, prev, lam, T_trans = self._state_now(batch, t)
        return lsa_maximize(beta_hat_now(beta, prev, float(lam), T_trans).float())


def _haal(args):
    raise NotImplementedError("haal_selector (look-ahead over time-interval sequences with deep-copied envs) is not built "
                              "yet (DESIGN.md, 'out of scope')")


REGISTRY = {"haa_selector": HAASelector, "haal_selector": _haal}
