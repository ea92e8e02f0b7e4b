"""Drop-in for the reference's src/flux/pipeline_tools.py (encode_images / prepare_text_input, pipeline_tools.py:7-52)."""
import torch
from torch import Tensor


def encode_images(pipeline, images: Tensor):
    """pipeline_tools.py:7-30: VAE-encode -> (x - shift) * scale -> _pack_latents -> ids.

    With a VAE attached (`pipeline.attach_vae(...)`) images go through the native encoder.  Already-encoded
    latents [B, 16, h, w] are accepted either way and go through the same native pack kernel and the same id
    construction, including the reference's diffusers-version fallback for the id grid size (pipeline_tools.py:22-29)."""
    if isinstance(images, torch.Tensor) and images.dim() == 4 and images.shape[1] == 16:
        latents = images.to(pipeline.device).to(pipeline.dtype)
    elif pipeline.vae is None:
        raise NotImplementedError("no VAE attached (pipeline.attach_vae): pass pre-encoded latents [B, 16, h, w]")
    else:
        images = pipeline.image_processor.preprocess(images)
        images = images.to(pipeline.device).to(pipeline.dtype)
        latents = pipeline.vae.encode(images).latent_dist.sample()
        latents = (latents - pipeline.vae.config.shift_factor) * pipeline.vae.config.scaling_factor
    tokens = pipeline._pack_latents(latents, *latents.shape)
    ids = pipeline._prepare_latent_image_ids(latents.shape[0], latents.shape[2], latents.shape[3], pipeline.device,
                                             pipeline.dtype)
    if tokens.shape[1] != ids.shape[0]:
        ids = pipeline._prepare_latent_image_ids(latents.shape[0], latents.shape[2] // 2, latents.shape[3] // 2,
                                                 pipeline.device, pipeline.dtype)
    return tokens, ids


def prepare_text_input(pipeline, prompts, max_sequence_length=512):
    """pipeline_tools.py:33-52: -> (prompt_embeds, pooled_prompt_embeds, text_ids); needs `pipeline.attach_text_encoders`."""
    return pipeline.encode_prompt(prompt=prompts, prompt_2=None, prompt_embeds=None, pooled_prompt_embeds=None,
                                  device=pipeline.device, num_images_per_prompt=1,
                                  max_sequence_length=max_sequence_length, lora_scale=None)
