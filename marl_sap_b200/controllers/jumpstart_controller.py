"""JumpstartMAC (/root/reference/src/controllers/jumpstart_controller.py:10-118): BasicMAC whose joint action is, with
probability ``jumpstart_epsilon(t_env)``, taken from a non-learning policy (``haa_selector``) instead of the agent
network + learning selector.  All shipped ``*_reda`` / ``*_iql`` / ``*_sap`` configs use it.  The Bernoulli draw is the
host's ``np.random.rand()`` per call, like the reference (:35), so one draw decides for the whole batch."""
import numpy as np

from ..action_selectors.non_rl_selectors import REGISTRY as non_rl_action_REGISTRY
from ..components.epsilon_schedules import DecayThenFlatSchedule
from .basic_controller import BasicMAC


class JumpstartMAC(BasicMAC):
    graph_capturable = False  # the HAA-vs-network decision is a host draw per step: a captured graph would freeze it

    def __init__(self, scheme, groups, args):
        super().__init__(scheme, groups, args)
        self.jumpstart_action_selector = non_rl_action_REGISTRY[args.jumpstart_action_selector](args)
        self.jumpstart_eps_schedule = DecayThenFlatSchedule(args.jumpstart_epsilon_start, args.jumpstart_epsilon_finish,
                                                            args.jumpstart_epsilon_anneal_time, decay="linear")
        self.jumpstart_epsilon = self.jumpstart_eps_schedule.eval(0)

    def select_actions(self, ep_batch, t_ep, t_env, bs=slice(None), test_mode=False):
        self.jumpstart_epsilon = self.jumpstart_eps_schedule.eval(t_env)
        if test_mode:
            self.jumpstart_epsilon = self.args.jumpstart_evaluation_epsilon
        if np.random.rand() < self.jumpstart_epsilon:
            return self.jumpstart_action_selector.select_action(ep_batch, t_ep)[bs]
        return super().select_actions(ep_batch, t_ep, t_env, bs=bs, test_mode=test_mode)
