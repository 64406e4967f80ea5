"""Learner registry with the reference's keys for the Q-learners of the SAP configs
(/root/reference/src/learners/__init__.py: ``mock_constellation_iql.yaml`` / ``iql_sap.yaml`` use ``q_learner``,
``filtered_iql.yaml`` uses ``filtered_q_learner``).  The policy-gradient, COMA, MADDPG and behaviour-cloning learners of the
reference are outside SURVEY.md section 8."""
from .q_learner import REGISTRY, FilteredQLearner, QLearner

__all__ = ["REGISTRY", "QLearner", "FilteredQLearner"]
