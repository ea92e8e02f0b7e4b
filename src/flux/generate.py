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
    config_path = config_path or os.environ.get("XFL_CONFIG")
    if not config_path:
        return {}
    with open(config_path, "r") as f:
        return yaml.safe_load(f)


def prepare_params(prompt=None, prompt_2=None, height: Optional[int] = 512, width: Optional[int] = 512,
                   num_inference_steps: int = 28, timesteps: List[int] = None, guidance_scale: float = 3.5,
                   num_images_per_prompt: Optional[int] = 1, generator=None, latents=None, prompt_embeds=None,
                   pooled_prompt_embeds=None, output_type: Optional[str] = "pil", return_dict: bool = True,
                   joint_attention_kwargs: Optional[Dict[str, Any]] = None, callback_on_step_end=None,
                   callback_on_step_end_tensor_inputs: List[str] = ["latents"], max_sequence_length: int = 512, **kwargs):
    return (prompt, prompt_2, height, width, num_inference_steps, timesteps, guidance_scale, num_images_per_prompt,
            generator, latents, prompt_embeds, pooled_prompt_embeds, output_type, return_dict, joint_attention_kwargs,
            callback_on_step_end, callback_on_step_end_tensor_inputs, max_sequence_length)


def seed_everything(seed: int = 42):
    torch.backends.cudnn.deterministic = True
    torch.manual_seed(seed)
    np.random.seed(seed)


def _signal(x, device):
    """generate.py:170-176 with D6: [C, L] gets a batch axis, [B, C, L] is kept; float32 on the device (D3/D4)."""
    if x is None:
        return None
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(np.asarray(x))
    if x.dim() == 2:
        x = x.unsqueeze(0)
    return x.to(device=device, dtype=torch.float32).contiguous()


@torch.no_grad()
def generate(model, pipeline, conditions: List[Condition] = None, config_path: str = None,
             model_config: Optional[Dict[str, Any]] = {}, condition_scale: float = 1.0, default_lora: bool = False,
             additional_condition1=None, additional_condition2=None, additional_condition3=None,
             additional_condition4=None, use_brain_condition: bool = True, fuse_flag: bool = True, **params):
    model_config = model_config or get_config(config_path).get("model", {})
    if condition_scale != 1:
        for name, module in pipeline.transformer.named_modules():
            if name.endswith(".attn"):
                module.c_factor = torch.ones(1, 1) * condition_scale
    self = pipeline
    (prompt, prompt_2, height, width, num_inference_steps, timesteps, guidance_scale, num_images_per_prompt, generator,
     latents, prompt_embeds, pooled_prompt_embeds, output_type, return_dict, joint_attention_kwargs, callback_on_step_end,
     callback_on_step_end_tensor_inputs, max_sequence_length) = prepare_params(**params)
    eeg_only_replace = bool(params.get("eeg_only_replace", False))

    height = height or self.default_sample_size * self.vae_scale_factor
    width = width or self.default_sample_size * self.vae_scale_factor
    self.check_inputs(prompt, prompt_2, height, width, prompt_embeds=prompt_embeds,
                      pooled_prompt_embeds=pooled_prompt_embeds,
                      callback_on_step_end_tensor_inputs=callback_on_step_end_tensor_inputs,
                      max_sequence_length=max_sequence_length)
    self._guidance_scale = guidance_scale
    self._joint_attention_kwargs = joint_attention_kwargs
    self._interrupt = False

    if prompt is not None and isinstance(prompt, str):
        batch_size = 1
    elif prompt is not None and isinstance(prompt, list):
        batch_size = len(prompt)
    else:
        batch_size = prompt_embeds.shape[0]
    device = self._execution_device
    lora_scale = self.joint_attention_kwargs.get("scale", None) if self.joint_attention_kwargs is not None else None
    self.transformer.set_lora_scale(1.0 if lora_scale is None else lora_scale)  # what scale_lora_layers does per forward
    prompt_embeds, pooled_prompt_embeds, text_ids = self.encode_prompt(
        prompt=prompt, prompt_2=prompt_2, prompt_embeds=prompt_embeds, pooled_prompt_embeds=pooled_prompt_embeds,
        device=device, num_images_per_prompt=num_images_per_prompt, max_sequence_length=max_sequence_length,
        lora_scale=lora_scale)

    # ---- neural-signal conditioning, once per call (generate.py:168-258)
    if use_brain_condition:
        eeg = _signal(additional_condition1, device)
        fnirs = _signal(additional_condition2, device)
        ppg = _signal(additional_condition3, device)
        motion = _signal(additional_condition4, device)
        if eeg is not None:
            eeg = model.spatial_pyramid_pooling(eeg, model.eeg_fixed_length)
        if fnirs is not None:
            fnirs = model.spatial_pyramid_pooling(fnirs, model.fnirs_fixed_length)
        if ppg is not None:
            ppg = model.spatial_pyramid_pooling(ppg, model.ppg_fixed_length)
        if motion is not None:
            motion = model.spatial_pyramid_pooling(motion, model.motion_fixed_length)

        prompt_embeds_brain = pooled_prompt_embeds_brain = None
        if eeg is not None:
            eeg_features = model.eeg_projection(eeg)
            prompt_embeds_brain = model.fuse_eeg(eeg_features, model.ppg_projection(ppg)) if ppg is not None else eeg_features
        if fnirs is not None:
            fnirs_features = model.fnirs_projection(fnirs)
            pooled_prompt_embeds_brain = (model.fuse_fnirs(fnirs_features, model.motion_projection(motion))
                                          if motion is not None else fnirs_features)

        if prompt_embeds_brain is not None and pooled_prompt_embeds_brain is not None:
            if fuse_flag:  # generate.py:240-255: DUAN(x = text embedding, c = brain embedding) replaces the embedding
                prompt_embeds = model.duan_norm_prompt(prompt_embeds, prompt_embeds_brain)
                pooled_prompt_embeds = model.duan_norm_pooled(pooled_prompt_embeds.unsqueeze(1),
                                                              pooled_prompt_embeds_brain.unsqueeze(1)).squeeze(1)
            else:  # generate.py:256-258
                prompt_embeds = model.to_model_dtype(prompt_embeds_brain)
                pooled_prompt_embeds = model.to_model_dtype(pooled_prompt_embeds_brain)
        elif eeg_only_replace and prompt_embeds_brain is not None:  # D5 opt-in
            prompt_embeds = model.to_model_dtype(prompt_embeds_brain)

    # ---- latents, condition tokens, ids (generate.py:260-287)
    num_channels_latents = self.transformer.config.in_channels // 4
    latents, latent_image_ids = self.prepare_latents(batch_size * num_images_per_prompt, num_channels_latents, height,
                                                     width, prompt_embeds.dtype, device, generator, latents)
    condition_latents = condition_ids = condition_type_ids = None
    use_condition = conditions is not None or []
    if use_condition:
        assert len(conditions) <= 1, "Only one condition is supported for now."
        if not default_lora:
            pipeline.set_adapters(conditions[0].condition_type)
        toks, ids_l, types = [], [], []
        for condition in conditions:
            tokens, ids, type_id = condition.encode(self)
            toks.append(tokens)
            ids_l.append(ids)
            types.append(type_id)
        condition_latents = torch.cat(toks, dim=1)
        condition_ids = torch.cat(ids_l, dim=0)
        condition_type_ids = torch.cat(types, dim=0)  # unused downstream, like the reference

    # ---- sigma schedule (generate.py:289-310)
    sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
    image_seq_len = latents.shape[1]
    mu = calculate_shift(image_seq_len, self.scheduler.config.base_image_seq_len, self.scheduler.config.max_image_seq_len,
                         self.scheduler.config.base_shift, self.scheduler.config.max_shift)
    timesteps, num_inference_steps = retrieve_timesteps(self.scheduler, num_inference_steps, device, timesteps, sigmas, mu=mu)
    num_warmup_steps = max(len(timesteps) - num_inference_steps * self.scheduler.order, 0)
    self._num_timesteps = len(timesteps)

    # ---- step-invariant DiT work, once (hoisted out of generate.py:313-345)
    B = latents.shape[0]
    T = len(timesteps)
    n_cond = condition_latents.shape[1] if use_condition else 0
    # cache_cond: with model_config.independent_condition the condition branch is step-invariant and runs once per edit
    plan = self.transformer.plan(B, prompt_embeds.shape[1], latents.shape[1], n_cond, T, model_config,
                                 self.transformer.c_factor(), cache_cond=True)
    plan.set_ids(text_ids, latent_image_ids, condition_ids if use_condition else None)
    step_t = [float(t) / 1000.0 for t in timesteps for _ in range(B)]  # the embedder sees sigma * 1000 (transformer.py:95)
    guidance = [float(guidance_scale)] * B if self.transformer.config.guidance_embeds else None
    plan.prepare(prompt_embeds, pooled_prompt_embeds, condition_latents if use_condition else None, step_t, guidance, c_t=0.0)

    latents = latents.to(torch.bfloat16).contiguous()
    noise_pred = torch.empty_like(latents)
    with self.progress_bar(total=num_inference_steps) as progress_bar:
        for i, t in enumerate(timesteps):
            if self.interrupt:
                continue
            plan.step(i, latents, noise_pred)
            latents = euler_step(latents, noise_pred, self.scheduler.advance())  # scheduler.step (generate.py:349)
            if callback_on_step_end is not None:
                callback_kwargs = {k: locals()[k] for k in callback_on_step_end_tensor_inputs}
                callback_outputs = callback_on_step_end(self, i, t, callback_kwargs)
                latents = callback_outputs.pop("latents", latents)
                if "prompt_embeds" in callback_outputs:
                    raise NotImplementedError("changing prompt_embeds mid-loop invalidates the prepared conditioning")
            if i == len(timesteps) - 1 or ((i + 1) > num_warmup_steps and (i + 1) % self.scheduler.order == 0):
                progress_bar.update()

    if output_type == "latent":
        image = latents
    else:
        if self.vae is None:
            raise NotImplementedError("no VAE in this build (SURVEY.md §8f.2): call generate(..., output_type='latent')")
        latents = self._unpack_latents(latents, height, width, self.vae_scale_factor)
        latents = (latents / self.vae.config.scaling_factor) + self.vae.config.shift_factor
        image = self.vae.decode(latents, return_dict=False)[0]
        image = self.image_processor.postprocess(image, output_type=output_type)
    self.maybe_free_model_hooks()

    if condition_scale != 1:
        for name, module in pipeline.transformer.named_modules():
            if name.endswith(".attn"):
                del module.c_factor
    if not return_dict:
        return (image,)
    return FluxPipelineOutput(images=image)
