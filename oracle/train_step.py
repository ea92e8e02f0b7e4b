"""ORACLE — test infrastructure only (see oracle/flux_dit.py header for who may import this).

Restatement of the training objective of the reference, `OminiModel.step` (/root/reference/src/train/model.py:569-729):
rectified-flow interpolation x_t = (1-t) x_0 + t x_1 with t = sigmoid(N(0,1)), neural conditioning in the *step* fuse
order (model.py:656-701), tranformer_forward with guidance = 1 (model.py:651-655, 705-723) and
loss = mse(pred, x_1 - x_0) (model.py:726).  Gradients come from torch autograd over this restatement.

Pinned bit-for-bit (loss) and to 1e-6 (LoRA gradients) against the reference's own `step` executed on the CPU through
oracle/ref_harness.py (tests/golden/ref_v1.npz 'step_*', tests/test_reference_pins_cpu.py).  The VAE / text encoders are
outside the path: `image` / `condition` are latents [B,16,h,w], text enters as embeddings.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import flux_dit as O
from . import sampler as OS


def lora_param_names(P: Dict[str, torch.Tensor]):
    return sorted(k for k in P if k.endswith(".lora_A.weight") or k.endswith(".lora_B.weight"))


def flow_step(P, cfg, batch: dict, model_config: Optional[dict] = None, conditioner=None, use_brain_condition: bool = False,
              fuse_flag: bool = True, dtype=torch.float32, generator: Optional[torch.Generator] = None):
    """-> (loss, aux) following model.py:569-729.  batch keys: image [B,16,h,w] latents, condition [B,16,h,w] latents,
    prompt_embeds [B,Nt,J], pooled_prompt_embeds [B,P], position_delta ([dy, dx],), optional position_scale, eeg / fnirs /
    ppg / motion [B,C,L]; optional 't' [B] and 'noise' (x_1) override the random draws."""
    imgs = batch["image"]
    dev = imgs.device
    x_0 = OS.pack_latents(imgs.to(dtype))  # encode_images (pipeline_tools.py:7-30) with the VAE factored out
    img_ids = OS.prepare_latent_image_ids(imgs.shape[2], imgs.shape[3]).to(dev)
    prompt_embeds, pooled = batch["prompt_embeds"].to(dtype), batch["pooled_prompt_embeds"].to(dtype)
    text_ids = torch.zeros(prompt_embeds.shape[1], 3, device=dev)
    B = imgs.shape[0]
    if "t" in batch:
        t = batch["t"].to(dev).float()
    else:
        t = torch.sigmoid(torch.randn((B,), device=dev, generator=generator))  # :590
    x_1 = batch["noise"].to(dtype) if "noise" in batch else torch.randn(x_0.shape, device=dev, dtype=x_0.dtype,
                                                                         generator=generator)  # :591
    t_ = t.unsqueeze(1).unsqueeze(1)
    x_t = ((1 - t_) * x_0 + t_ * x_1).to(dtype)  # :592-593
    cond = batch["condition"]
    condition_latents = OS.pack_latents(cond.to(dtype))
    delta = batch["position_delta"][0]
    scale = float(batch.get("position_scale", [1.0])[0])
    condition_ids = OS.condition_ids(OS.prepare_latent_image_ids(cond.shape[2], cond.shape[3]).to(dev), delta, scale)
    guidance = torch.ones_like(t) if cfg.guidance_embeds else None  # :651-655
    if use_brain_condition:  # :656-701 ("step" fuse order, D2)
        prompt_embeds, pooled = conditioner.conditioning(prompt_embeds, pooled, batch.get("eeg"), batch.get("fnirs"),
                                                         batch.get("ppg"), batch.get("motion"), fuse_flag=fuse_flag,
                                                         mode="step")
    pred = O.tranformer_forward(P, cfg, condition_latents, condition_ids, None, model_config or {}, 0, hidden_states=x_t,
                                encoder_hidden_states=prompt_embeds, pooled_projections=pooled, timestep=t.to(dtype),
                                img_ids=img_ids, txt_ids=text_ids, guidance=guidance)
    loss = F.mse_loss(pred, (x_1 - x_0), reduction="mean")  # :726
    return loss, dict(pred=pred, t=t, x_0=x_0, x_1=x_1, x_t=x_t, prompt_embeds=prompt_embeds, pooled=pooled)


def flow_step_grads(P, cfg, batch, **kw):
    """loss and d loss / d (every LoRA factor), fp32 autograd."""
    names = lora_param_names(P)
    Pg = dict(P)
    for n in names:
        Pg[n] = P[n].detach().clone().requires_grad_(True)
    loss, aux = flow_step(Pg, cfg, batch, **kw)
    grads = torch.autograd.grad(loss, [Pg[n] for n in names], allow_unused=True)
    return loss.detach(), {n: (g if g is not None else torch.zeros_like(Pg[n])) for n, g in zip(names, grads)}, aux
