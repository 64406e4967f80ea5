"""Independent Q-learning on the device-resident episode buffer.

``QLearner``          <- /root/reference/src/learners/q_learner.py:10-231
``FilteredQLearner``  <- /root/reference/src/learners/filtered_q_learner.py:10-260

Same constructor ``(mac, scheme, logger, args)``, ``train(batch, t_env, episode_num)``, ``cuda()``, ``save_models`` /
``load_models``, target updates (hard every ``target_update_interval_or_tau`` training steps when > 1, Polyak otherwise),
``standardise_rewards`` / ``standardise_returns``, double Q, Adam, gradient clipping and the same ``logger.log_stat`` keys
(loss, grad_norm, td_error_abs, q_taken_mean, target_mean, avg_num_conflicts, avg_beta).  Built for batches that already
live on the GPU:

* the agent is evaluated for ALL time steps of the sampled episodes in one pass when it is feed-forward
  (``use_rnn=False``: [B (T+1) n, obs] rows through three GEMMs) instead of T + 1 small passes; recurrent agents are
  unrolled over time as in the reference;
* the two diagnostics are device array ops (``utils.rollout_stats``), not Python triple loops with ``.item()`` per action;
* the filtered learner takes the top-M tasks from ``sap_topm_from_beta`` (stable order, the same rule as the env and the
  selector) instead of ``th.topk`` on the host copy;
* with ``torch.distributed`` initialised (one process per GPU, each with its own replay shard) the gradients are averaged
  with one flat all-reduce over NCCL before clipping, and the standardisation statistics are combined over the ranks, so
  all replicas stay identical.

Mixers (``vdn`` / ``qmix``) need ``batch["state"]``, which no SAP scheme defines: rejected."""
from __future__ import annotations

import copy

import torch as th
from torch.optim import Adam

from .. import _lib
from ..components.standardize_stream import RunningMeanStd
from ..utils import rollout_stats
from ..utils.dist import all_reduce_gradients

_MASKED = -9999999.0


class QLearner:
    def __init__(self, mac, scheme, logger, args):
        if getattr(args, "mixer", None) is not None:
            raise ValueError("Mixer {} needs batch['state'], which the SAP schemes do not define".format(args.mixer))
        self.args, self.mac, self.logger = args, mac, logger
        self.params = list(mac.parameters())
        self.mixer = None
        self.optimiser = Adam(params=self.params, lr=args.lr)
        self.target_mac = copy.deepcopy(mac)
        self.training_steps = 0
        self.last_target_update_step = 0
        self.last_target_update_episode = 0
        self.log_stats_t = -self.args.learner_log_interval - 1
        self.n, self.m = args.n, args.m
        device = "cuda" if getattr(args, "use_cuda", True) else "cpu"
        if args.standardise_returns:
            self.ret_ms = RunningMeanStd(shape=(self.n,), device=device)
        if args.standardise_rewards:
            self.rew_ms = RunningMeanStd(shape=(self.n,), device=device)

    # ------------------------------------------------------------------ Q-values of every time step
    def _agent_outputs(self, mac, batch):
        """[B, T+1, n, A]: ``mac.forward(batch, t)`` for every t (q_learner.py:68-74)."""
        B, T1 = batch.batch_size, batch.max_seq_length
        if not getattr(self.args, "use_rnn", False) and getattr(batch, "agent_in", None) is None:
            # feed-forward agent: time is just more rows
            rows = [mac._build_inputs(batch, t) for t in range(T1)]
            x = th.stack(rows, dim=0)                              # [T+1, B n, in]
            q, _ = mac.agent(x.reshape(T1 * B * mac.n, -1), mac.agent.init_hidden().expand(T1 * B * mac.n, -1))
            if mac.agent_output_type == "pi_logits":
                q = th.softmax(q, dim=-1)
            return q.view(T1, B, mac.n, -1).transpose(0, 1)
        mac.init_hidden(B)
        return th.stack([mac.forward(batch, t=t) for t in range(T1)], dim=1)

    def _q_values(self, outs, batch, top):
        """Per-task Q-values [B, T', n, m] from the agent outputs (identity for the plain learner)."""
        return outs

    def _top_tasks(self, batch):
        return None

    # ------------------------------------------------------------------ one gradient step (q_learner.py:54-149)
    def train(self, batch, t_env, episode_num):
        rewards = batch["rewards"][:, :-1].float()
        actions = batch["actions"][:, :-1].to(th.int64)
        terminated = batch["terminated"][:, :-1].float()
        mask = batch["filled"][:, :-1].float()
        mask[:, 1:] = mask[:, 1:] * (1 - terminated[:, :-1])
        avail = batch["avail_actions"]
        if self.args.standardise_rewards:
            self.rew_ms.update(rewards)
            rewards = (rewards - self.rew_ms.mean) / th.sqrt(self.rew_ms.var)

        top = self._top_tasks(batch)
        q_all = self._q_values(self._agent_outputs(self.mac, batch), batch, top)              # [B, T+1, n, m]
        chosen = th.gather(q_all[:, :-1], dim=3, index=actions).squeeze(3)
        with th.no_grad():
            tq = self._q_values(self._agent_outputs(self.target_mac, batch), batch, top)[:, 1:]
            tq = tq.masked_fill(avail[:, 1:] == 0, _MASKED)
            if self.args.double_q:
                live = q_all.detach()[:, 1:].masked_fill(avail[:, 1:] == 0, _MASKED)
                target_max = th.gather(tq, 3, live.max(dim=3, keepdim=True)[1]).squeeze(3)
            else:
                target_max = tq.max(dim=3)[0]
            if self.args.standardise_returns:
                target_max = target_max * th.sqrt(self.ret_ms.var) + self.ret_ms.mean
            targets = rewards + self.args.gamma * (1 - terminated) * target_max
            if self.args.standardise_returns:
                self.ret_ms.update(targets)
                targets = (targets - self.ret_ms.mean) / th.sqrt(self.ret_ms.var)

        mask = mask.expand_as(chosen)
        masked_td = (chosen - targets) * mask
        loss = (masked_td ** 2).sum() / mask.sum()

        self.optimiser.zero_grad()
        loss.backward()
        all_reduce_gradients(self.params)   # replicas on other GPUs: one flat NCCL all-reduce (no-op for one process)
        grad_norm = th.nn.utils.clip_grad_norm_(self.params, self.args.grad_norm_clip)
        self.optimiser.step()

        self.training_steps += 1
        tau = self.args.target_update_interval_or_tau
        if tau > 1 and (self.training_steps - self.last_target_update_step) / tau >= 1.0:
            self._update_targets_hard()
            self.last_target_update_step = self.training_steps
        elif tau <= 1.0:
            self._update_targets_soft(tau)
        if not getattr(self.args, "use_mps_action_selection", True):
            self.mac.update_action_selector_agent()

        if t_env - self.log_stats_t >= self.args.learner_log_interval:
            elems = mask.sum().item()
            self.logger.log_stat("loss", loss.item(), t_env)
            self.logger.log_stat("grad_norm", float(grad_norm), t_env)
            self.logger.log_stat("td_error_abs", masked_td.abs().sum().item() / elems, t_env)
            self.logger.log_stat("q_taken_mean", (chosen * mask).sum().item() / (elems * self.args.n), t_env)
            self.logger.log_stat("target_mean", (targets * mask).sum().item() / (elems * self.args.n), t_env)
            self.log_stats_t = t_env
            self.logger.log_stat("avg_num_conflicts", self.calc_conflicting_actions(actions), t_env)
            self.logger.log_stat("avg_beta", self.calc_raw_benefits(batch["beta"], actions), t_env)
        return loss.detach()

    # ------------------------------------------------------------------ diagnostics (q_learner.py:157-191), on the device
    def calc_conflicting_actions(self, actions):
        return rollout_stats.calc_conflicting_actions(actions, self.m)

    def calc_raw_benefits(self, beta, actions):
        return rollout_stats.calc_raw_benefits(beta, actions)

    # ------------------------------------------------------------------ targets / checkpoints (q_learner.py:193-231)
    def _update_targets_hard(self):
        self.target_mac.load_state(self.mac)

    def _update_targets_soft(self, tau):
        with th.no_grad():
            for tp, p in zip(self.target_mac.parameters(), self.mac.parameters()):
                tp.mul_(1.0 - tau).add_(p, alpha=tau)

    def cuda(self):
        self.mac.cuda()
        self.target_mac.cuda()

    def save_models(self, path):
        self.mac.save_models(path)
        th.save(self.optimiser.state_dict(), "{}/opt.th".format(path))

    def load_models(self, path):
        self.mac.load_models(path)
        self.target_mac.load_models(path)   # like the reference: the target network is not checkpointed separately
        self.optimiser.load_state_dict(th.load("{}/opt.th".format(path), map_location=lambda storage, loc: storage))


class FilteredQLearner(QLearner):
    """The agent emits M + 1 values per agent: Q for its top-M tasks and one baseline for every other task
    (filtered_q_learner.py:67-92, 100-111)."""

    def _top_tasks(self, batch):
        beta = batch["beta"]
        B, T1 = beta.shape[0], beta.shape[1]
        M = self.args.env_args["M"]
        flat = beta.reshape(B * T1, *beta.shape[2:]).contiguous()
        if flat.is_cuda and flat.dtype in (th.float16, th.float32):
            n, m = flat.shape[1], flat.shape[2]
            L = flat.shape[3] if flat.dim() == 4 else 1
            top = th.empty(B * T1, n, M, dtype=th.int32, device=flat.device)
            _lib.check(_lib.load().sap_topm_from_beta(flat.data_ptr(), _lib.sap_dtype(flat.dtype), B * T1, n, m, L, M,
                                                      top.data_ptr(), _lib.stream_ptr(flat.device)), "sap_topm_from_beta")
            return top.view(B, T1, n, M).long()
        total = flat.float().sum(-1) if flat.dim() == 4 else flat.float()
        return th.topk(total, k=M, dim=-1).indices.view(B, T1, -1, M)

    def _q_values(self, outs, batch, top):
        base = outs[..., -1:].expand(-1, -1, -1, self.args.m)
        return base.scatter(3, top[:, :outs.shape[1]], outs[..., :-1])


REGISTRY = {"q_learner": QLearner, "filtered_q_learner": FilteredQLearner}
