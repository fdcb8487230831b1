"""Timestep respacing (reference ``diffusion/respace.py``).

``space_timesteps`` is pure integer/float64 host arithmetic and must match the
reference bit-exactly (including Python's banker's rounding in ``round``).
``SpacedDiffusion`` rebuilds the beta schedule on the retained timesteps; the
index remap the reference performs with a fresh host->device copy on every
model call (``respace.py:129-134``) is a device-resident lookup table here,
and for the native denoiser it is folded into the time-embedding table.
"""
import numpy as np
import torch as th

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Set of original-process timesteps to keep (reference respace.py:8-61).

    ``"ddimN"``: the first integer stride that yields exactly N steps.
    list / ``"a,b,c"``: split the process into len(counts) equal sections and take
    ``count`` evenly spaced (float stride, Python ``round``) steps from each.
    """
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                picked = range(0, num_timesteps, stride)
                if len(picked) == want:
                    return set(picked)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    kept, start = [], 0
    for sec, count in enumerate(section_counts):
        size = base + (1 if sec < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


class SpacedDiffusion(GaussianDiffusion):
    """Diffusion over a subset of the base process' timesteps (reference respace.py:64-115)."""

    def __init__(self, use_timesteps, **kwargs):
        base_betas = np.asarray(kwargs["betas"], dtype=np.float64)
        self.original_num_steps = len(base_betas)
        self.use_timesteps = set(use_timesteps)
        # retained steps in increasing order; beta'_j = 1 - abar[i_j] / abar[i_{j-1}] with abar[i_{-1}] = 1.  Same float64
        # operations, element by element, as the reference's loop over base_diffusion.alphas_cumprod (respace.py:73-87).
        kept = np.array(sorted(t for t in self.use_timesteps if 0 <= t < self.original_num_steps), dtype=np.int64)
        abar = np.cumprod(1.0 - base_betas, axis=0)[kept]
        abar_before = np.concatenate(([1.0], abar[:-1]))
        self.timestep_map = [int(t) for t in kept]
        kwargs["betas"] = 1 - abar / abar_before
        super().__init__(**kwargs)
        self._map_dev = {}

    def p_mean_variance(self, model, *args, **kwargs):  # pylint: disable=signature-differs
        return super().p_mean_variance(self._wrap_model(model), *args, **kwargs)

    def condition_mean(self, cond_fn, *args, **kwargs):
        return super().condition_mean(self._wrap_model(cond_fn), *args, **kwargs)

    def few_shot_style_finetune_losses(self, model, *args, **kwargs):
        return super().few_shot_style_finetune_losses(self._wrap_model(model), *args, **kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps, owner=self)

    def _scale_timesteps(self, t):
        # scaling is done by the wrapped model
        return t

    # -- native-path hooks -----------------------------------------------------------
    def _model_time_index(self, t_index):
        if self.rescale_timesteps:
            raise NotImplementedError("rescale_timesteps=True feeds float timesteps to the model; the reference "
                                      "never enables it (utils/model_util.py:177)")
        return self.timestep_map[t_index]

    def map_tensor(self, device, dtype=th.long):
        key = (str(device), dtype)
        m = self._map_dev.get(key)
        if m is None:
            m = th.tensor(self.timestep_map, device=device, dtype=dtype)
            self._map_dev[key] = m
        return m

    def _map_model_t(self, t):
        new_ts = self.map_tensor(t.device, t.dtype)[t]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return new_ts


class _WrappedModel:
    """Callable that remaps respaced indices to original timesteps before calling the model."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps, owner=None):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps
        self._owner = owner
        self._cache = {}

    def _map(self, ts):
        if self._owner is not None:
            return self._owner.map_tensor(ts.device, ts.dtype)
        key = (str(ts.device), ts.dtype)
        if key not in self._cache:
            self._cache[key] = th.tensor(self.timestep_map, device=ts.device, dtype=ts.dtype)
        return self._cache[key]

    def __call__(self, x, ts, **kwargs):
        new_ts = self._map(ts)[ts]
        if self.rescale_timesteps:
            new_ts = new_ts.float() * (1000.0 / self.original_num_steps)
        return self.model(x, new_ts, **kwargs)
