"""Drop-in for the reference's src/flux/generate.py: `generate()` (generate.py:72-394), `get_config`, `seed_everything`.

Same signature, same attribute mutations on the pipeline, same return types.  What differs is where the work happens:

  * neural conditioning (generate.py:168-258) runs on the native fp32 CS3 / DGF kernels through the model's
    `eeg_projection` / `fuse_eeg` / `duan_norm_prompt` ... attributes (loongx_b200/cs3.py);
  * everything in the DiT that does not depend on the latents (context embedding, condition embedding, RoPE tables,
    temb and the AdaLN modulation of all blocks for ALL timesteps of the sigma schedule) is prepared once per call, and the
    loop body is lx_dit_step + lx_euler_step.

Documented deviations from the literal reference: D1 signals are passed to the encoders as [B, C, L]
(the reference's `.flatten(1)` crashes in EEGEncoder.forward); D3/D4 the conditioning runs in float32; D5 EEG-only is a
no-op unless `eeg_only_replace=True` is passed; D6 batched [B, C, L] signals are accepted.
"""
import os
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import yaml

from loongx_b200.dit import euler_step
from loongx_b200.pipeline import FluxPipelineOutput
from loongx_b200.sampler import calculate_shift, retrieve_timesteps

from .condition import Condition


def get_config(config_path: str = None):
    """generate.py:16-23: the YAML named by `config_path` or $XFL_CONFIG ({} when neither is set)."""
    path = config_path or os.environ.get("XFL_CONFIG")
    if not path:
        return {}
    with open(path, "r") as fh:
        return yaml.safe_load(fh)


# keyword parameters of FluxPipeline.__call__ that generate() honours (generate.py:25-65), with their defaults
_CALL_DEFAULTS = dict(prompt=None, prompt_2=None, height=512, width=512, num_inference_steps=28, timesteps=None,
                      guidance_scale=3.5, num_images_per_prompt=1, generator=None, latents=None, prompt_embeds=None,
                      pooled_prompt_embeds=None, output_type="pil", return_dict=True, joint_attention_kwargs=None,
                      callback_on_step_end=None, callback_on_step_end_tensor_inputs=("latents",), max_sequence_length=512)


def prepare_params(prompt=None, prompt_2=None, height: Optional[int] = 512, width: Optional[int] = 512,
                   num_inference_steps: int = 28, timesteps: List[int] = None, guidance_scale: float = 3.5,
                   num_images_per_prompt: Optional[int] = 1, generator=None, latents=None, prompt_embeds=None,
                   pooled_prompt_embeds=None, output_type: Optional[str] = "pil", return_dict: bool = True,
                   joint_attention_kwargs: Optional[Dict[str, Any]] = None, callback_on_step_end=None,
                   callback_on_step_end_tensor_inputs: List[str] = ["latents"], max_sequence_length: int = 512, **kwargs):
    """Same signature and positional result as the reference's helper (generate.py:25-65); unknown kwargs are dropped."""
    given = locals()
    return tuple(given[name] for name in _CALL_DEFAULTS)


def seed_everything(seed: int = 42):
    torch.backends.cudnn.deterministic = True
    for seeder in (torch.manual_seed, np.random.seed):
        seeder(seed)


def _signal(x, device):
    """generate.py:170-176 with D6: [C, L] gets a batch axis, [B, C, L] is kept; float32 on the device (D3/D4)."""
    if x is None:
        return None
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(np.asarray(x))
    if x.dim() == 2:
        x = x.unsqueeze(0)
    return x.to(device=device, dtype=torch.float32).contiguous()


def _set_c_factor(transformer, value):
    """generate.py:90-94 / 385-389: `condition_scale` travels as an attribute on every `*.attn` module."""
    for name, module in transformer.named_modules():
        if name.endswith(".attn"):
            if value is None:
                del module.c_factor
            else:
                module.c_factor = torch.ones(1, 1) * value


def _neural_conditioning(model, prompt_embeds, pooled, raw_signals, device, fuse_flag, eeg_only_replace):
    """generate.py:168-258, once per call: CS3 encoders on the length-normalised signals, DGF fusion of the modality
    pairs, then either the DUAN fuse with the text embeddings (fuse_flag) or their replacement."""
    lengths = (model.eeg_fixed_length, model.fnirs_fixed_length, model.ppg_fixed_length, model.motion_fixed_length)
    sig = []
    for raw, n in zip(raw_signals, lengths):
        x = _signal(raw, device)
        sig.append(model.spatial_pyramid_pooling(x, n) if x is not None else None)
    eeg, fnirs, ppg, motion = sig
    tokens_b = pooled_b = None
    if eeg is not None:
        tokens_b = model.eeg_projection(eeg)
        if ppg is not None:
            tokens_b = model.fuse_eeg(tokens_b, model.ppg_projection(ppg))
    if fnirs is not None:
        pooled_b = model.fnirs_projection(fnirs)
        if motion is not None:
            pooled_b = model.fuse_fnirs(pooled_b, model.motion_projection(motion))
    if tokens_b is not None and pooled_b is not None:
        if fuse_flag:  # :240-255: DUAN(x = text embedding, c = brain embedding) replaces the embedding
            fused_pooled = model.duan_norm_pooled(pooled.unsqueeze(1), pooled_b.unsqueeze(1)).squeeze(1)
            return model.duan_norm_prompt(prompt_embeds, tokens_b), fused_pooled
        return model.to_model_dtype(tokens_b), model.to_model_dtype(pooled_b)  # :256-258
    if eeg_only_replace and tokens_b is not None:  # D5 opt-in
        return model.to_model_dtype(tokens_b), pooled
    return prompt_embeds, pooled


def _encode_conditions(pipe, conditions, default_lora):
    """generate.py:273-287 -> (tokens [B, N, 64], ids [N, 3]) of the (single) condition, or (None, None)."""
    if not (conditions is not None or []):
        return None, None
    assert len(conditions) <= 1, "Only one condition is supported for now."
    if not default_lora:
        pipe.set_adapters(conditions[0].condition_type)
    encoded = [c.encode(pipe) for c in conditions]  # (tokens, ids, type ids); the type ids are unused like in the reference
    return torch.cat([e[0] for e in encoded], dim=1), torch.cat([e[1] for e in encoded], dim=0)


@torch.no_grad()
def generate(model, pipeline, conditions: List[Condition] = None, config_path: str = None,
             model_config: Optional[Dict[str, Any]] = {}, condition_scale: float = 1.0, default_lora: bool = False,
             additional_condition1=None, additional_condition2=None, additional_condition3=None,
             additional_condition4=None, use_brain_condition: bool = True, fuse_flag: bool = True, **params):
    model_config = model_config or get_config(config_path).get("model", {})
    pipe = pipeline
    if condition_scale != 1:
        _set_c_factor(pipe.transformer, condition_scale)
    a = dict(zip(_CALL_DEFAULTS, prepare_params(**params)))
    height = a["height"] or pipe.default_sample_size * pipe.vae_scale_factor
    width = a["width"] or pipe.default_sample_size * pipe.vae_scale_factor
    pipe.check_inputs(a["prompt"], a["prompt_2"], height, width, prompt_embeds=a["prompt_embeds"],
                      pooled_prompt_embeds=a["pooled_prompt_embeds"],
                      callback_on_step_end_tensor_inputs=a["callback_on_step_end_tensor_inputs"],
                      max_sequence_length=a["max_sequence_length"])
    pipe._guidance_scale, pipe._joint_attention_kwargs, pipe._interrupt = a["guidance_scale"], a["joint_attention_kwargs"], False
    prompt = a["prompt"]
    batch_size = 1 if isinstance(prompt, str) else len(prompt) if isinstance(prompt, list) else a["prompt_embeds"].shape[0]
    device = pipe._execution_device
    jak = pipe.joint_attention_kwargs
    lora_scale = jak.get("scale", None) if jak is not None else None
    pipe.transformer.set_lora_scale(1.0 if lora_scale is None else lora_scale)  # what scale_lora_layers does per forward
    prompt_embeds, pooled_prompt_embeds, text_ids = pipe.encode_prompt(
        prompt=prompt, prompt_2=a["prompt_2"], prompt_embeds=a["prompt_embeds"], pooled_prompt_embeds=a["pooled_prompt_embeds"],
        device=device, num_images_per_prompt=a["num_images_per_prompt"], max_sequence_length=a["max_sequence_length"],
        lora_scale=lora_scale)
    if use_brain_condition:
        prompt_embeds, pooled_prompt_embeds = _neural_conditioning(
            model, prompt_embeds, pooled_prompt_embeds,
            (additional_condition1, additional_condition2, additional_condition3, additional_condition4), device, fuse_flag,
            bool(params.get("eeg_only_replace", False)))

    # latents, condition tokens, ids (generate.py:260-287)
    latents, latent_image_ids = pipe.prepare_latents(batch_size * a["num_images_per_prompt"],
                                                     pipe.transformer.config.in_channels // 4, height, width,
                                                     prompt_embeds.dtype, device, a["generator"], a["latents"])
    condition_latents, condition_ids = _encode_conditions(pipe, conditions, default_lora)

    # sigma schedule (generate.py:289-310)
    n_steps = a["num_inference_steps"]
    sc = pipe.scheduler.config
    mu = calculate_shift(latents.shape[1], sc.base_image_seq_len, sc.max_image_seq_len, sc.base_shift, sc.max_shift)
    timesteps, n_steps = retrieve_timesteps(pipe.scheduler, n_steps, device, a["timesteps"],
                                            np.linspace(1.0, 1 / n_steps, n_steps), mu=mu)
    num_warmup_steps = max(len(timesteps) - n_steps * pipe.scheduler.order, 0)
    pipe._num_timesteps = len(timesteps)

    # step-invariant DiT work, once (hoisted out of generate.py:313-345); cache_cond: with
    # model_config.independent_condition the condition branch is step-invariant too and runs once per edit
    B, T = latents.shape[0], len(timesteps)
    n_cond = condition_latents.shape[1] if condition_latents is not None else 0
    plan = pipe.transformer.plan(B, prompt_embeds.shape[1], latents.shape[1], n_cond, T, model_config,
                                 pipe.transformer.c_factor(), cache_cond=True)
    plan.set_ids(text_ids, latent_image_ids, condition_ids)
    step_t = [float(t) / 1000.0 for t in timesteps for _ in range(B)]  # the embedder sees sigma * 1000 (transformer.py:95)
    guidance = [float(a["guidance_scale"])] * B if pipe.transformer.config.guidance_embeds else None
    plan.prepare(prompt_embeds, pooled_prompt_embeds, condition_latents, step_t, guidance, c_t=0.0)

    latents = latents.to(torch.bfloat16).contiguous()
    noise_pred = torch.empty_like(latents)
    on_step_end = a["callback_on_step_end"]
    with pipe.progress_bar(total=n_steps) as progress_bar:
        for i, t in enumerate(timesteps):
            if pipe.interrupt:
                continue
            plan.step(i, latents, noise_pred)
            latents = euler_step(latents, noise_pred, pipe.scheduler.advance())  # scheduler.step (generate.py:349)
            if on_step_end is not None:
                visible = dict(latents=latents, prompt_embeds=prompt_embeds, pooled_prompt_embeds=pooled_prompt_embeds,
                               noise_pred=noise_pred, timestep=t)
                out = on_step_end(pipe, i, t, {k: visible[k] for k in a["callback_on_step_end_tensor_inputs"]})
                latents = out.pop("latents", latents)
                if "prompt_embeds" in out:
                    raise NotImplementedError("changing prompt_embeds mid-loop invalidates the prepared conditioning")
            if i == len(timesteps) - 1 or ((i + 1) > num_warmup_steps and (i + 1) % pipe.scheduler.order == 0):
                progress_bar.update()

    if a["output_type"] == "latent":
        image = latents
    else:
        if pipe.vae is None:
            raise NotImplementedError("no VAE attached (pipeline.attach_vae, SURVEY.md §8f.2): call "
                                      "generate(..., output_type='latent')")
        z = pipe._unpack_latents(latents, height, width, pipe.vae_scale_factor).float()  # fp32 like the reference's VAE
        z = z / pipe.vae.config.scaling_factor + pipe.vae.config.shift_factor
        image = pipe.image_processor.postprocess(pipe.vae.decode(z, return_dict=False)[0], output_type=a["output_type"])
    pipe.maybe_free_model_hooks()
    if condition_scale != 1:
        _set_c_factor(pipe.transformer, None)
    return FluxPipelineOutput(images=image) if a["return_dict"] else (image,)
