"""Shared-parameter multi-agent controller: glue between the episode batch, the torch agent and the kernel selector.

Keeps the interface of /root/reference/src/controllers/basic_controller.py:7-101 (``select_actions``, ``forward``,
``init_hidden``, ``parameters``, ``load_state``, ``cuda``, ``save_models``, ``load_models``,
``update_action_selector_agent``, ``_build_inputs``, ``_get_input_shape``; attributes ``agent``, ``action_selector``,
``hidden_states``), with two differences that matter on a GPU:

  * one device-resident agent serves training and action selection (the reference keeps a second CPU
    "selector_agent" unless ``use_mps_action_selection`` is set, :69-75); ``update_action_selector_agent`` is a no-op;
  * when the runner provides ``batch.agent_in`` (fp32 staging rows the env kernel already filled with
    ``float(obs[:, t])``), the per-step ``obs[:, t].float()`` / ``cat`` of :77-92 is skipped and only the optional
    last-action / agent-id columns are written.
"""
import torch as th

from ..action_selectors import REGISTRY as action_REGISTRY
from ..modules.agents import REGISTRY as agent_REGISTRY


def _obs_width(scheme):
    v = scheme["obs"]["vshape"]
    return v if isinstance(v, int) else v[0]


class BasicMAC:
    # select_actions is torch modules + device-side kernels only (no host decision per step): the runner may capture
    # its T-step loop in a CUDA graph.  Subclasses that decide on the host (JumpstartMAC) set this to False.
    graph_capturable = True

    def __init__(self, scheme, groups, args):
        self.args = args
        self.n = args.n
        self.agent_output_type = args.agent_output_type
        self.agent = agent_REGISTRY[args.agent](self._get_input_shape(scheme), args)
        self.selector_agent = self.agent
        self.action_selector = action_REGISTRY[args.action_selector](args)
        self.hidden_states = None
        self._filtered = type(self.action_selector).__name__.startswith("Filtered")

    # ------------------------------------------------------------------ acting
    def select_actions(self, ep_batch, t_ep, t_env, bs=slice(None), test_mode=False):
        q = self.forward(ep_batch, t_ep, test_mode=test_mode, action_selection_mode=True)[bs]
        avail = ep_batch["avail_actions"][:, t_ep][bs]
        if not self._filtered:
            return self.action_selector.select_action(q, avail, t_env, test_mode=test_mode, beta=None)
        top = getattr(ep_batch, "top_agent_tasks", None)  # the env's own top-M task indices (computed once)
        if top is not None:
            return self.action_selector.select_action(q, avail, t_env, test_mode=test_mode, top=top[bs])
        return self.action_selector.select_action(q, avail, t_env, test_mode=test_mode, beta=ep_batch["beta"][bs, t_ep])

    def supports_select_and_step(self, env, ep_batch):
        """True when ``select_and_step`` applies: the classic epsilon-greedy selector on Q-values and an env whose step
        kernel can select its own actions (``sap_rollout_step``)."""
        return (type(self.action_selector).__name__ == "EpsilonGreedyActionSelector" and self.agent_output_type == "q"
                and hasattr(self.action_selector, "fused_select_args") and hasattr(env, "supports_select_step")
                and env.supports_select_step(ep_batch))

    def select_and_step(self, ep_batch, t_ep, t_env, env, test_mode=False, agent_in=None):
        """``select_actions`` followed by ``env.step`` (episode_runner.py:75-84) as ONE kernel launch after the agent
        forward: the env's CTA draws / arg-maxes its own agents' actions and steps.  Returns the actions [B, n]."""
        q = self.forward(ep_batch, t_ep, test_mode=test_mode, action_selection_mode=True)
        sel, actions, keep = self.action_selector.fused_select_args(q, t_env, test_mode=test_mode)
        env.step_select(sel, actions, ep_batch, agent_in=agent_in)
        del keep
        return actions

    def forward(self, ep_batch, t, test_mode=False, action_selection_mode=False):
        outs, self.hidden_states = self.agent(self._build_inputs(ep_batch, t), self.hidden_states)
        if self.agent_output_type == "pi_logits":
            outs = th.softmax(outs, dim=-1)
        return outs.view(ep_batch.batch_size, self.n, -1)

    def init_hidden(self, batch_size):
        h0 = self.agent.init_hidden()  # [1, hidden]
        self.hidden_states = h0.unsqueeze(0).expand(batch_size, self.n, -1)

    # ------------------------------------------------------------------ agent inputs
    def _get_input_shape(self, scheme):
        width = _obs_width(scheme)
        if self.args.obs_last_action:
            width += scheme["actions_onehot"]["vshape"][0]
        if self.args.obs_agent_id:
            width += self.n
        return width

    def _build_inputs(self, batch, t):
        staged = getattr(batch, "agent_in", None)
        if staged is not None and getattr(batch, "agent_in_t", None) == t:
            return self._finish_staged_inputs(batch, t, staged)
        rows = batch.batch_size * self.n
        parts = [batch["obs"][:, t].float().reshape(rows, -1)]
        if self.args.obs_last_action:
            prev = batch["actions_onehot"][:, t - 1] if t > 0 else th.zeros_like(batch["actions_onehot"][:, 0])
            parts.append(prev.reshape(rows, -1))
        if self.args.obs_agent_id:
            eye = th.eye(self.n, device=batch.device)
            parts.append(eye.unsqueeze(0).expand(batch.batch_size, -1, -1).reshape(rows, -1))
        return parts[0] if len(parts) == 1 else th.cat(parts, dim=1)

    def _finish_staged_inputs(self, batch, t, staged):
        """Columns [0, obs) already hold float(obs[:, t]) (written by the env kernel); fill the optional rest."""
        col = _obs_width(batch.scheme)
        if self.args.obs_last_action:
            m = batch.scheme["actions_onehot"]["vshape"][0]
            block = staged[:, :, col:col + m]
            block.zero_()
            if t > 0:
                block.scatter_(2, batch["actions"][:, t - 1].long(), 1.0)
            col += m
        if self.args.obs_agent_id:  # written once per staging buffer (the runner may alternate between two)
            done = getattr(batch, "_agent_in_ids_done", None)
            if done is None:
                done = batch._agent_in_ids_done = set()
            if staged.data_ptr() not in done:
                staged[:, :, col:col + self.n] = th.eye(self.n, device=staged.device, dtype=staged.dtype)
                done.add(staged.data_ptr())
        return staged.view(batch.batch_size * self.n, -1)

    # ------------------------------------------------------------------ parameters / checkpoints (agent.th, :62-67)
    def parameters(self):
        return self.agent.parameters()

    def load_state(self, other_mac):
        self.agent.load_state_dict(other_mac.agent.state_dict())

    def cuda(self):
        self.agent.cuda()

    def save_models(self, path):
        th.save(self.agent.state_dict(), "{}/agent.th".format(path))

    def load_models(self, path):
        state = th.load("{}/agent.th".format(path), map_location=lambda storage, loc: storage)
        self.agent.load_state_dict(state)
        self.update_action_selector_agent()

    def update_action_selector_agent(self):
        return None  # single agent: nothing to synchronise
