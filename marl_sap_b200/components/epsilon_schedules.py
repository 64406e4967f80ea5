"""Exploration schedule used by the selectors.

Same constructor, attributes (``start, finish, time_length, delta, decay, exp_scaling``) and values as
/root/reference/src/components/epsilon_schedules.py:4-25:
  linear: eps(T) = max(finish, start - (start - finish) / time_length * T)
  exp:    eps(T) = clip(exp(-T / s), finish, start) with s = -time_length / ln(finish)  (s = 1 when finish <= 0)
"""
import math


def _linear(sched, T):
    return max(sched.finish, sched.start - sched.delta * T)


def _exponential(sched, T):
    return min(sched.start, max(sched.finish, math.exp(-T / sched.exp_scaling)))


_DECAYS = {"linear": _linear, "exp": _exponential}


class DecayThenFlatSchedule:
    def __init__(self, start, finish, time_length, decay="exp"):
        self.start, self.finish, self.time_length, self.decay = start, finish, time_length, decay
        self.delta = (start - finish) / time_length
        if decay == "exp":
            self.exp_scaling = -time_length / math.log(finish) if finish > 0 else 1

    def eval(self, T):
        fn = _DECAYS.get(self.decay)
        return None if fn is None else fn(self, T)
