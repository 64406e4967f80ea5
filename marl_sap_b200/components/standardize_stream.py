"""Running mean / variance of a stream of batches (/root/reference/src/components/standarize_stream.py:9-45, used by the
learners for ``standardise_rewards`` / ``standardise_returns``), on the device.

Same update rule (Chan et al.'s parallel combination of (count, mean, M2)) and the same attributes ``mean``, ``var``,
``count``.  With ``torch.distributed`` initialised the batch moments are first combined over the ranks (one all-reduce of
[count, sum, sum of squares]), so that every replica standardises with the statistics of the whole job."""
from __future__ import annotations

import torch as th
import torch.distributed as dist


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=(), device="cpu"):
        self.mean = th.zeros(shape, dtype=th.float32, device=device)
        self.var = th.ones(shape, dtype=th.float32, device=device)
        self.count = epsilon

    def update(self, arr):
        rows = arr.reshape(-1, arr.size(-1)).float()
        count = rows.shape[0]
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            packed = th.cat([rows.sum(0), (rows * rows).sum(0), rows.new_tensor([float(count)])])
            dist.all_reduce(packed)
            d = rows.shape[1]
            count = int(round(packed[-1].item()))
            mean = packed[:d] / count
            var = (packed[d:2 * d] - count * mean * mean) / max(count - 1, 1)   # unbiased, like torch.var
        else:
            mean, var = rows.mean(0), rows.var(0)
        self.update_from_moments(mean, var, count)

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        total = self.count + batch_count
        delta = batch_mean - self.mean
        m2 = self.var * self.count + batch_var * batch_count + delta * delta * (self.count * batch_count / total)
        self.mean = self.mean + delta * (batch_count / total)
        self.var = m2 / total
        self.count = total
