"""ORACLE — test infrastructure only (see oracle/flux_dit.py header for who may import this).

Restatement of the sampler side of the reference: the denoise loop of src/flux/generate.py:260-372 and the diffusers
0.31.0 helpers it calls (`FlowMatchEulerDiscreteScheduler.set_timesteps/step`, `calculate_shift`,
`FluxPipeline._pack_latents / _unpack_latents / _prepare_latent_image_ids`), plus the id
arithmetic of src/flux/condition.py:126-137.  The loop and the id arithmetic are pinned bit-for-bit to the reference's
own generate.py / condition.py / pipeline_tools.py executed on the CPU (oracle/ref_harness.py, tests/golden/ref_v1.npz);
diffusers itself is absent from /root/reference and from this image; its sigma schedule is pinned to Black Forest Labs'
own `get_schedule` (torchtitan's copy, tests/test_oracle_cpu.py::test_oracle_sigma_schedule_vs_bfl_get_schedule), the
packing / id helpers by the closed-form checks in tests/test_oracle_cpu.py.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import flux_dit as O


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.16):
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def flow_match_sigmas(num_inference_steps: int, image_seq_len: int, base_image_seq_len=256, max_image_seq_len=4096,
                      base_shift=0.5, max_shift=1.15) -> torch.Tensor:
    """generate.py:290-306 + set_timesteps(sigmas=..., mu=...): returns sigmas [n+1] float32 (last = 0)."""
    sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
    mu = calculate_shift(image_seq_len, base_image_seq_len, max_image_seq_len, base_shift, max_shift)
    sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1) ** 1.0)
    sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32)
    return torch.cat([sigmas, torch.zeros(1)])


def pack_latents(latents: torch.Tensor) -> torch.Tensor:
    B, C, h, w = latents.shape
    latents = latents.view(B, C, h // 2, 2, w // 2, 2).permute(0, 2, 4, 1, 3, 5)
    return latents.reshape(B, (h // 2) * (w // 2), C * 4)


def unpack_latents(latents: torch.Tensor, height: int, width: int, vae_scale_factor: int = 16) -> torch.Tensor:
    B, _, ch = latents.shape
    height, width = height // vae_scale_factor, width // vae_scale_factor
    latents = latents.view(B, height, width, ch // 4, 2, 2).permute(0, 3, 1, 4, 2, 5)
    return latents.reshape(B, ch // 4, height * 2, width * 2)


def prepare_latent_image_ids(height: int, width: int, dtype=torch.float32) -> torch.Tensor:
    """_prepare_latent_image_ids(batch, height, width, ...) of diffusers 0.31.0: height/width are the latent sizes."""
    ids = torch.zeros(height // 2, width // 2, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(height // 2)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(width // 2)[None, :]
    return ids.reshape((height // 2) * (width // 2), 3).to(dtype)


def condition_ids(ids: torch.Tensor, position_delta=None, position_scale: float = 1.0) -> torch.Tensor:
    """condition.py:126-137 on a copy."""
    ids = ids.clone()
    if position_delta is not None:
        ids[:, 1] += position_delta[0]
        ids[:, 2] += position_delta[1]
    if position_scale != 1.0:
        scale_bias = (position_scale - 1.0) / 2
        ids[:, 1] *= position_scale
        ids[:, 2] *= position_scale
        ids[:, 1] += scale_bias
        ids[:, 2] += scale_bias
    return ids


@torch.no_grad()
def denoise(P, cfg, latents, prompt_embeds, pooled, text_ids, img_ids, cond_latents=None, cond_ids=None,
            num_inference_steps: int = 28, guidance_scale: float = 3.5, model_config: Optional[dict] = None,
            c_factor: Optional[float] = None, t_in_model_dtype: bool = False):
    """generate.py:313-369: n x (tranformer_forward + scheduler.step) -> final packed latents.

    t_in_model_dtype=True reproduces the reference's `t.expand(B).to(latents.dtype)` (generate.py:319), which under
    bf16 quantises the timestep; False keeps float32 timesteps (what the fp32 reference configuration computes)."""
    sigmas = flow_match_sigmas(num_inference_steps, latents.shape[1])
    B = latents.shape[0]
    dev = latents.device
    for i in range(num_inference_steps):
        t = sigmas[i] * 1000.0
        timestep = t.expand(B).to(dev)
        timestep = timestep.to(latents.dtype) if t_in_model_dtype else timestep.float()
        guidance = torch.tensor([guidance_scale], device=dev).expand(B) if cfg.guidance_embeds else None
        noise_pred = O.tranformer_forward(
            P, cfg, cond_latents, cond_ids, None, model_config or {}, 0, hidden_states=latents,
            encoder_hidden_states=prompt_embeds, pooled_projections=pooled, timestep=(timestep / 1000).to(latents.dtype)
            if t_in_model_dtype else timestep / 1000, img_ids=img_ids, txt_ids=text_ids, guidance=guidance, c_factor=c_factor)
        x = latents.to(torch.float32)
        x = x + (sigmas[i + 1] - sigmas[i]).to(dev) * noise_pred.to(torch.float32)
        latents = x.to(noise_pred.dtype)
    return latents
