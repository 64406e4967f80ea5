"""Non-learning jump-start policies (/root/reference/src/action_selectors/non_rl_selectors.py).

HAASelector ("haa_selector", :10-50): every env takes the optimal assignment of its CURRENT handover-aware benefits,
``linear_sum_assignment(beta_hat(beta, prev_assigns)[..., 0], maximize=True)``.  The reference loops over the batch on
the host (``env.beta_hat`` + scipy per env); here the whole batch is a few device array ops for beta_hat (the formula of
real_constellation_env.py:282-328 / mock_constellation_env.py:228-274) and one launch of the batched assignment kernel.
Like the reference it works on the ``beta`` / ``prev_assigns`` state fields of the episode batch, i.e. on benefits
already rounded to the scheme dtype; when the runner has bound its batched env (``bind_env``) the same values are
taken from the env's planes instead of materialising the lazy ``beta`` field.

HAALSelector ("haal_selector", :54-145): look-ahead over every way to cut the next L steps into intervals.  The
reference deep-copies the env per sequence and steps the copies on the host; here the "copies" are three small device
tensors per sequence (step counter, prev_assigns, accumulated reward) for ALL envs of the batch at once, the rewards of a
simulated step are a dozen array ops on the benefit planes (the reward rule of real_constellation_env.py:135-160), and every
interval's assignment is one launch of the batched assignment kernel.
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


def time_interval_sequences(L):
    """utils/methods.py:309-349: every split of steps 0 .. L-1 into consecutive intervals, in the reference's order."""
    out = []

    def rec(seq, start):
        if start == L:
            out.append(tuple(seq))
            return
        for end in range(start, L):
            rec(seq + [(start, end)], end + 1)

    rec([], 0)
    return out


class HAALSelector:
    """non_rl_selectors.py:54-145 on the runner's batched real env (``bind_env``).  The reference asserts the episode
    runner because it deep-copies one live env per batch element; the batched version has no such limit."""

    def __init__(self, args):
        self.args = args
        self.envs = None
        self._env = None

    def bind_env(self, env):
        self._env = env

    @staticmethod
    def _window(env, k):
        """beta of step k in float64 [B, n, m, L] (real_constellation_env.py:167-170), straight from the planes."""
        L = env.L
        win = env.planes[:, k:k + L].double()
        if win.shape[1] < L:
            win = th.cat([win, win.new_zeros(win.shape[0], L - win.shape[1], env.n, env.m)], dim=1)
        beta = win.permute(0, 2, 3, 1)
        if env.task_prios is not None:
            beta = beta * env.task_prios.double().view(1, 1, -1, 1)
        return beta.expand(env.B, -1, -1, -1) if beta.shape[0] == 1 and env.B > 1 else beta

    @staticmethod
    def _beta_hat(env, beta, prev):
        out = beta.clone()
        out[..., 0] = beta_hat_now(beta, prev, env.lambda_, env.T_trans)
        return out

    @staticmethod
    def _step_reward(env, beta, prev, a):
        """sum_i reward_i of one env step (:145-160): conflict counts, beta_hat at the chosen entries, split if positive."""
        B, n, m = env.B, env.n, env.m
        cnt = th.zeros(B, m, dtype=th.float64, device=a.device).scatter_add_(1, a, th.ones(B, n, dtype=th.float64, device=a.device))
        chosen = beta.gather(2, a.view(B, n, 1, 1).expand(-1, -1, 1, beta.shape[-1])).squeeze(2)   # [B, n, L]
        if env.T_trans is None:
            pen = (a != prev).double()
        else:
            pen = env.T_trans.double()[prev, a]
        bh = chosen[..., 0] - env.lambda_ * pen * (chosen.sum(-1) > 1e-12).double()
        r = th.where(bh > 0, bh / cnt.gather(1, a), bh)
        return r.sum(1)

    def select_action(self, batch=None, t=None):
        env = self._env
        if env is None or env.kind != "real":
            raise RuntimeError("haal_selector needs the runner's batched real env (runner.setup binds it)")
        k0 = env.t_host
        eff = min(env.L, env.T - k0)
        best_val, best = None, None
        for tis in time_interval_sequences(eff):
            k, prev = k0, env.prev.long()
            val = th.zeros(env.B, dtype=th.float64, device=env.device)
            first = None
            for i, (t0, t1) in enumerate(tis):
                total = self._beta_hat(env, self._window(env, k), prev).sum(-1)
                a = lsa_maximize(total.float().contiguous())
                for _ in range(t1 - t0 + 1):
                    val = val + self._step_reward(env, self._window(env, k), prev, a)
                    k, prev = k + 1, a
                if i == 0:
                    first = a
            if best is None:
                best_val, best = val, first
            else:
                better = val > best_val                       # strict: the first sequence wins ties (:109-112)
                best_val = th.where(better, val, best_val)
                best = th.where(better.unsqueeze(1), first, best)
        return best


REGISTRY = {"haa_selector": HAASelector, "haal_selector": HAALSelector}
