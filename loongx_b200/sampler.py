"""Host-side scheduler logic of the denoise loop (generate.py:289-309, 349), restated from diffusers 0.31.0
`FlowMatchEulerDiscreteScheduler` / `calculate_shift` / `retrieve_timesteps` as recorded in SURVEY.md App. A.6.

Pure scalar arithmetic on a few dozen floats (numpy float32 where the reference uses float32 tensors); the per-step
tensor update itself is the native lx_euler_step kernel.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np


def calculate_shift(image_seq_len: int, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.16) -> float:
    """diffusers.pipelines.flux.pipeline_flux.calculate_shift: linear interpolation of mu in the image sequence length."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


@dataclass
class SchedulerConfig:
    """FLUX.1-dev scheduler/scheduler_config.json."""

    num_train_timesteps: int = 1000
    shift: float = 3.0
    use_dynamic_shifting: bool = True
    base_shift: float = 0.5
    max_shift: float = 1.15
    base_image_seq_len: int = 256
    max_image_seq_len: int = 4096


class FlowMatchEulerDiscreteScheduler:
    """The subset generate() touches: config, order, set_timesteps(sigmas=, mu=), timesteps, sigmas, step()."""

    order = 1

    def __init__(self, config: Optional[SchedulerConfig] = None):
        self.config = config or SchedulerConfig()
        self.timesteps: np.ndarray = np.zeros((0,), dtype=np.float32)
        self.sigmas: np.ndarray = np.zeros((1,), dtype=np.float32)
        self._step_index: Optional[int] = None
        self.num_inference_steps = 0

    @staticmethod
    def time_shift(mu: float, sigma: float, t: np.ndarray) -> np.ndarray:
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None,
                      sigmas: Optional[Sequence[float]] = None, mu: Optional[float] = None) -> None:
        cfg = self.config
        if cfg.use_dynamic_shifting and mu is None:
            raise ValueError("you have to pass a value for `mu` when `use_dynamic_shifting` is set to be `True`")
        if sigmas is None:
            ts = np.linspace(cfg.num_train_timesteps, 1, num_inference_steps)  # sigma_max = 1, sigma_min = 1/1000
            sigmas = ts / cfg.num_train_timesteps
        sigmas = np.asarray(sigmas, dtype=np.float64)
        self.num_inference_steps = len(sigmas)
        if cfg.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = cfg.shift * sigmas / (1 + (cfg.shift - 1) * sigmas)
        sigmas = sigmas.astype(np.float32)  # torch.from_numpy(sigmas).to(dtype=torch.float32)
        self.timesteps = sigmas * np.float32(cfg.num_train_timesteps)
        self.sigmas = np.concatenate([sigmas, np.zeros(1, dtype=np.float32)])
        self._step_index = None

    @property
    def step_index(self) -> Optional[int]:
        return self._step_index

    def dt(self, i: int) -> float:
        """(sigma_{i+1} - sigma_i) as the float32 difference the reference computes."""
        return float(np.float32(self.sigmas[i + 1]) - np.float32(self.sigmas[i]))

    def advance(self) -> float:
        """dt of the current step; increments the internal step index like scheduler.step()."""
        if self._step_index is None:
            self._step_index = 0
        d = self.dt(self._step_index)
        self._step_index += 1
        return d


def retrieve_timesteps(scheduler: FlowMatchEulerDiscreteScheduler, num_inference_steps: Optional[int] = None, device=None,
                       timesteps: Optional[List[int]] = None, sigmas: Optional[Sequence[float]] = None, **kwargs):
    """diffusers retrieve_timesteps for this scheduler: custom `timesteps` are not supported by
    FlowMatchEulerDiscreteScheduler.set_timesteps (the reference would raise ValueError too)."""
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed.")
    if timesteps is not None:
        raise ValueError("FlowMatchEulerDiscreteScheduler.set_timesteps does not support custom timestep schedules")
    scheduler.set_timesteps(num_inference_steps, device=device, sigmas=sigmas, **kwargs)
    return scheduler.timesteps, len(scheduler.timesteps)


def latent_image_ids(height: int, width: int) -> np.ndarray:
    """FluxPipeline._prepare_latent_image_ids(batch, height, width): [height*width, 3] = (0, row, col), integers."""
    ids = np.zeros((height, width, 3), dtype=np.float32)
    ids[..., 1] += np.arange(height, dtype=np.float32)[:, None]
    ids[..., 2] += np.arange(width, dtype=np.float32)[None, :]
    return ids.reshape(height * width, 3)


def shard_range(n_items: int, rank: int, world: int):
    """Reference list sharding (inference.py:126-128): contiguous chunks of len//world, last rank takes the rest."""
    chunk = n_items // world
    start = rank * chunk
    end = n_items if rank == world - 1 else start + chunk
    return start, end
