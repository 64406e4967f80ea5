"""Minimal stat logger with the interface the runners use (/root/reference/src/utils/logging.py:5-73:
``log_stat(key, value, t)`` and an in-memory ``stats`` dict); tensorboard / wandb / sacred sinks are out of scope."""
import logging
from collections import defaultdict


class Logger:
    def __init__(self, console_logger=None):
        self.console_logger = console_logger if console_logger is not None else logging.getLogger("marl_sap_b200")
        self.stats = defaultdict(lambda: [])

    def log_stat(self, key, value, t, to_sacred=True):
        self.stats[key].append((t, value))

    def print_recent_stats(self):
        items = ["{}: {:.4f}".format(k, float(v[-1][1])) for k, v in sorted(self.stats.items())]
        self.console_logger.info("Recent Stats | " + " | ".join(items))
