"""Shared-parameter multi-agent controller (mirror of /root/reference/src/controllers/basic_controller.py:7-101).

Glue only: build inputs -> torch agent forward -> kernel selector.  A single on-device agent is used for both
training and selection (the reference's CPU "selector_agent" copy, :69-75, is what
``use_mps_action_selection=True`` disables); ``update_action_selector_agent`` is kept as a no-op.
"""
import torch as th

from ..action_selectors import REGISTRY as action_REGISTRY
from ..modules.agents import REGISTRY as agent_REGISTRY


class BasicMAC:
    def __init__(self, scheme, groups, args):
        self.n = args.n
        self.args = args
        input_shape = self._get_input_shape(scheme)
        self._build_agents(input_shape)
        self.agent_output_type = args.agent_output_type
        self.action_selector = action_REGISTRY[args.action_selector](args)
        self.hidden_states = None

    def select_actions(self, ep_batch, t_ep, t_env, bs=slice(None), test_mode=False):
        agent_outputs = self.forward(ep_batch, t_ep, test_mode=test_mode, action_selection_mode=True)
        avail_actions = ep_batch["avail_actions"][:, t_ep]
        top = getattr(ep_batch, "top_agent_tasks", None)
        if top is not None and hasattr(self.action_selector, "select_action") and \
                self.action_selector.__class__.__name__.startswith("Filtered"):
            return self.action_selector.select_action(agent_outputs[bs], avail_actions[bs], t_env, test_mode=test_mode,
                                                      top=top[bs])
        beta = ep_batch["beta"][bs, t_ep] if self.action_selector.__class__.__name__.startswith("Filtered") else None
        return self.action_selector.select_action(agent_outputs[bs], avail_actions[bs], t_env, test_mode=test_mode, beta=beta)

    def forward(self, ep_batch, t, test_mode=False, action_selection_mode=False):
        agent_inputs = self._build_inputs(ep_batch, t)
        agent_outs, self.hidden_states = self.agent(agent_inputs, self.hidden_states)
        if self.agent_output_type == "pi_logits":
            agent_outs = th.nn.functional.softmax(agent_outs, dim=-1)
        return agent_outs.view(ep_batch.batch_size, self.n, -1)

    def init_hidden(self, batch_size):
        self.hidden_states = self.agent.init_hidden().unsqueeze(0).expand(batch_size, self.n, -1)

    def parameters(self):
        return self.agent.parameters()

    def load_state(self, other_mac):
        self.agent.load_state_dict(other_mac.agent.state_dict())

    def cuda(self):
        self.agent.cuda()

    def save_models(self, path):
        th.save(self.agent.state_dict(), "{}/agent.th".format(path))

    def load_models(self, path):
        self.agent.load_state_dict(th.load("{}/agent.th".format(path), map_location=lambda storage, loc: storage))
        self.update_action_selector_agent()

    def _build_agents(self, input_shape):
        self.agent = agent_REGISTRY[self.args.agent](input_shape, self.args)
        self.selector_agent = self.agent  # one device-resident agent serves both roles

    def update_action_selector_agent(self):
        return None

    def _build_inputs(self, batch, t):
        bs = batch.batch_size
        staged = getattr(batch, "agent_in", None)
        if staged is not None and getattr(batch, "agent_in_t", None) == t:
            # the env kernel already wrote float(obs[:, t]) into columns [0, obs_size) of the staging rows
            o = batch.scheme["obs"]["vshape"]
            o = o if isinstance(o, int) else o[0]
            if self.args.obs_last_action:
                m = batch.scheme["actions_onehot"]["vshape"][0]
                staged[:, :, o:o + m] = 0
                if t > 0:
                    staged[:, :, o:o + m].scatter_(2, batch["actions"][:, t - 1].long(), 1.0)
                o += m
            if self.args.obs_agent_id and getattr(batch, "agent_in_eye_t", None) is None:
                staged[:, :, o:o + self.n] = th.eye(self.n, device=staged.device)
                batch.agent_in_eye_t = True
            return staged.view(bs * self.n, -1)
        inputs = [batch["obs"][:, t].float()]
        if self.args.obs_last_action:
            if t == 0:
                inputs.append(th.zeros_like(batch["actions_onehot"][:, t]))
            else:
                inputs.append(batch["actions_onehot"][:, t - 1])
        if self.args.obs_agent_id:
            inputs.append(th.eye(self.n, device=batch.device).unsqueeze(0).expand(bs, -1, -1))
        if len(inputs) == 1:
            return inputs[0].reshape(bs * self.n, -1)
        return th.cat([x.reshape(bs * self.n, -1) for x in inputs], dim=1)

    def _get_input_shape(self, scheme):
        input_shape = scheme["obs"]["vshape"]
        if self.args.obs_last_action:
            input_shape += scheme["actions_onehot"]["vshape"][0]
        if self.args.obs_agent_id:
            input_shape += self.n
        return input_shape
