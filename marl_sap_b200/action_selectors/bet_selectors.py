"""ContinuousActionSelector ("continuous", /root/reference/src/action_selectors/bet_selectors.py:4-24): the agent outputs a
bid per task and the "action" is that bid vector plus Gaussian noise whose standard deviation follows the epsilon
schedule; the env (``bids_as_actions``) turns the bid matrix of every env into an assignment (linear_sum_assignment,
here the batched ``sap_lsa_maximize``).  Noise: ``torch.randn`` on the device or ``inject_draws(z=...)``."""
from __future__ import annotations

import torch as th

from .selectors import _KernelSelectorBase


class ContinuousActionSelector(_KernelSelectorBase):
    def __init__(self, args):
        super().__init__(args)
        self.variance = self.schedule.eval(0)

    def select_action(self, agent_inputs, avail_actions, t_env, test_mode=False, state=None, beta=None):
        x = agent_inputs.detach()
        if getattr(self.args, "softmax_agent_inputs", False):
            x = th.softmax(x, dim=1)  # :13 (dim=1, as written there)
        self.variance = self.args.evaluation_epsilon if test_mode else self.schedule.eval(t_env)  # :15-20
        inj = self._injected or {}
        self._injected = None
        z = inj.get("z")
        if z is None:
            z = th.randn(x.shape, device=x.device, dtype=x.dtype)
        return x + z.to(x.device, x.dtype) * float(self.variance)  # th.normal(mean, std)

    def action_log_prob(self, actions, old_agent_inputs):
        return th.distributions.Normal(old_agent_inputs, self.variance).log_prob(actions)  # :23-24
