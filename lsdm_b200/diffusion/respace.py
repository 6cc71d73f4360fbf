"""``space_timesteps`` / ``SpacedDiffusion`` -- mirror of the reference's ``diffusion/respace.py``.

The reference wraps the model in ``_WrappedModel`` which computes the mapped timesteps and then passes the
RAW (compact) ``ts`` to the model (respace.py:125-130); the schedule tables are indexed by the same compact
``t``.  Here that is simply "one ``t`` for both", which is what the C ABI takes.
"""
from __future__ import annotations

import numpy as np

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Reference respace.py:8-61: the retained timesteps for ``'ddimN'`` or per-section counts."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == desired:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    start, steps = 0, []
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return set(steps)


class _WrappedModel:
    """Reference respace.py:117-130.  Kept so ``diffusion._wrap_model(model).model`` works; the model sees raw ``ts``."""

    def __init__(self, model, timestep_map, rescale_timesteps, original_num_steps):
        self.model = model
        self.timestep_map = timestep_map
        self.rescale_timesteps = rescale_timesteps
        self.original_num_steps = original_num_steps

    def __call__(self, x_t, mask, ts, given_objs, given_cats, y):
        return self.model(x_t, mask, ts, given_objs, given_cats, y)


class SpacedDiffusion(GaussianDiffusion):
    """Reference respace.py:64-115: keeps ``use_timesteps`` of a base process, re-deriving the betas."""

    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(kwargs["betas"])
        abar = np.cumprod(1.0 - np.array(kwargs["betas"], dtype=np.float64))
        last, new_betas = 1.0, []
        for i, a in enumerate(abar):
            if i in self.use_timesteps:
                new_betas.append(1 - a / last)
                last = a
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(new_betas)
        super().__init__(**kwargs)

    def _wrap_model(self, model):
        if isinstance(model, _WrappedModel):
            return model
        return _WrappedModel(model, self.timestep_map, self.rescale_timesteps, self.original_num_steps)

    def _scale_timesteps(self, t):
        return t
