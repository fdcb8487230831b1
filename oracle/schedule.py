"""float64 / integer restatement of the schedule math (oracle; test infrastructure only)."""
import math

import numpy as np


def cosine_betas(n, max_beta=0.999):
    """reference diffusion/gaussian_diffusion.py:40-44, :49-66"""
    def abar(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2
    return np.array([min(1 - abar((i + 1) / n) / abar(i / n), max_beta) for i in range(n)], dtype=np.float64)


def linear_betas(n, scale_betas=1.0):
    """reference diffusion/gaussian_diffusion.py:31-39"""
    s = scale_betas * 1000 / n
    return np.linspace(s * 0.0001, s * 0.02, n, dtype=np.float64)


def named_betas(name, n):
    if name == "cosine":
        return cosine_betas(n)
    if name == "linear":
        return linear_betas(n)
    raise NotImplementedError(name)


def space_timesteps(num_timesteps, section_counts):
    """reference diffusion/respace.py:8-61 (banker's rounding of Python's round is part of the contract)"""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            k = int(section_counts[4:])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == k:
                    return sorted(set(range(0, num_timesteps, i)))
            raise ValueError("no integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per, extra = num_timesteps // len(section_counts), num_timesteps % len(section_counts)
    start, out = 0, []
    for i, c in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < c:
            raise ValueError("section too small")
        frac = 1 if c <= 1 else (size - 1) / (c - 1)
        cur = 0.0
        for _ in range(c):
            out.append(start + round(cur))
            cur += frac
        start += size
    return sorted(set(out))


class Schedule:
    """All per-timestep tables of a (respaced) process in float64.
    reference: GaussianDiffusion.__init__ gaussian_diffusion.py:183-219; SpacedDiffusion.__init__ respace.py:73-87"""

    def __init__(self, base_betas, use_timesteps=None, sigma_small=True):
        base_betas = np.asarray(base_betas, dtype=np.float64)
        if use_timesteps is None:
            use_timesteps = range(len(base_betas))
        use = set(use_timesteps)
        base_abar = np.cumprod(1.0 - base_betas, axis=0)
        last, betas, tmap = 1.0, [], []
        for i, a in enumerate(base_abar):
            if i in use:
                betas.append(1 - a / last)
                last = a
                tmap.append(i)
        self.timestep_map = tmap
        b = np.array(betas, dtype=np.float64)
        self.betas = b
        self.N = len(b)
        al = 1.0 - b
        self.abar = np.cumprod(al, axis=0)
        self.abar_prev = np.append(1.0, self.abar[:-1])
        self.sqrt_abar = np.sqrt(self.abar)
        self.sqrt_1m_abar = np.sqrt(1.0 - self.abar)
        self.sqrt_recip_abar = np.sqrt(1.0 / self.abar)
        self.sqrt_recipm1_abar = np.sqrt(1.0 / self.abar - 1)
        self.post_var = b * (1.0 - self.abar_prev) / (1.0 - self.abar)
        self.post_logvar_clipped = np.log(np.append(self.post_var[1], self.post_var[1:]))
        self.coef1 = b * np.sqrt(self.abar_prev) / (1.0 - self.abar)
        self.coef2 = (1.0 - self.abar_prev) * np.sqrt(al) / (1.0 - self.abar)
        if sigma_small:   # FIXED_SMALL, gaussian_diffusion.py:375-378
            self.var, self.logvar = self.post_var, self.post_logvar_clipped
        else:             # FIXED_LARGE, :371-374
            self.var = np.append(self.post_var[1], b[1:])
            self.logvar = np.log(self.var)
