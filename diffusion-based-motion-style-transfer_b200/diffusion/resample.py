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
        """(timesteps int64 [batch], importance weights fp32 [batch]).  ``data_range`` restricts the support to those
        steps (the finetune loop passes range(num_timesteps - skip), training_loop.py:242-244).  One
        ``np.random.choice`` call with the normalised weights, exactly as the reference draws them (:42-63), so a seeded
        ``np.random`` reproduces its timesteps."""
        weights_all = np.asarray(self.weights(), dtype=np.float64)
        if data_range is None:
            support = len(weights_all)
            prob = weights_all / weights_all.sum()
        else:
            support = data_range
            restricted = weights_all[data_range]
            prob = restricted / restricted.sum()
        drawn = np.random.choice(support, size=(batch_size,), p=prob)
        importance = 1 / (len(prob) * prob[drawn])
        if th.device(device).type != "cuda":
            return th.from_numpy(drawn).long().to(device), th.from_numpy(importance).float().to(device)
        # a blocking host-to-device copy drains the stream once per training step; stage the few bytes in pinned memory
        # and copy asynchronously instead (ring of staging buffers: a slot is reused only after its copy has completed)
        ring = self.__dict__.setdefault("_h2d_ring", {})
        key = (batch_size, str(device))
        if key not in ring:
            ring[key] = [0, [(th.empty(batch_size, dtype=th.int64).pin_memory(), th.empty(batch_size, dtype=th.float32).pin_memory(),
                              th.cuda.Event()) for _ in range(4)], [False] * 4]
        state = ring[key]
        k = state[0] % 4
        state[0] += 1
        t_host, w_host, ev = state[1][k]
        if state[2][k]:
            ev.synchronize()
        t_host.copy_(th.from_numpy(np.asarray(drawn, dtype=np.int64)))
        w_host.copy_(th.from_numpy(np.asarray(importance, dtype=np.float32)))
        with th.cuda.device(device):
            timesteps = t_host.to(device, non_blocking=True)
            weights = w_host.to(device, non_blocking=True)
            ev.record()
        state[2][k] = True
        return timesteps, weights


class UniformSampler(ScheduleSampler):
    def __init__(self, diffusion):
        self.diffusion = diffusion
        self._weights = np.ones([diffusion.num_timesteps])

    def weights(self):
        return self._weights


class LossAwareSampler(ScheduleSampler):
    """Marker base class the loop tests for (training_loop.py:272); no loss-aware sampler is configured."""
