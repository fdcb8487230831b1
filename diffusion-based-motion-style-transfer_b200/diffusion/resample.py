"""Timestep samplers of the training loop (reference ``diffusion/resample.py:8-72``).

Only the uniform sampler is on the finetune path (``train/training_loop.py:94-95``); sampling is a host-side
``numpy.random.choice`` exactly as in the reference, so a seeded ``np.random`` reproduces its draws.
"""
from abc import ABC, abstractmethod

import numpy as np
import torch as th


def create_named_schedule_sampler(name, diffusion):
    if name == "uniform":
        return UniformSampler(diffusion)
    raise NotImplementedError(f"unknown schedule sampler: {name} (only 'uniform' is used by the finetune loop)")


class ScheduleSampler(ABC):
    @abstractmethod
    def weights(self):
        """numpy array of positive weights, one per diffusion step"""

    def sample(self, batch_size, device, data_range=None):
        """(timesteps int64 [batch], importance weights fp32 [batch]); ``data_range`` restricts the support
        (the finetune loop passes range(num_timesteps - skip), training_loop.py:242-244)."""
        w = self.weights()
        p = w / np.sum(w)
        if data_range is None:
            indices_np = np.random.choice(len(p), size=(batch_size,), p=p)
        else:
            w_1 = self.weights()[data_range]
            p = w_1 / np.sum(w_1)
            indices_np = np.random.choice(data_range, size=(batch_size,), p=p)
        indices = th.from_numpy(indices_np).long().to(device)
        weights_np = 1 / (len(p) * p[indices_np])
        weights = th.from_numpy(weights_np).float().to(device)
        return indices, weights


class UniformSampler(ScheduleSampler):
    def __init__(self, diffusion):
        self.diffusion = diffusion
        self._weights = np.ones([diffusion.num_timesteps])

    def weights(self):
        return self._weights


class LossAwareSampler(ScheduleSampler):
    """Marker base class the loop tests for (training_loop.py:272); no loss-aware sampler is configured."""
