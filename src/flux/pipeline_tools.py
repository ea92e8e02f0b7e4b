"""Drop-in for the reference's src/flux/pipeline_tools.py (encode_images / prepare_text_input, pipeline_tools.py:7-52)."""
import torch
from torch import Tensor


def _to_latents(pipeline, images) -> Tensor:
    """Pictures -> scaled VAE latents [B, 16, h, w] (pipeline_tools.py:8-13).  Tensors that already are latents (16
    channels) skip the VAE; without a VAE only those are accepted."""
    already_encoded = isinstance(images, Tensor) and images.dim() == 4 and images.shape[1] == 16
    if already_encoded:
        return images.to(pipeline.device).to(pipeline.dtype)
    vae = pipeline.vae
    if vae is None:
        raise NotImplementedError("no VAE attached (pipeline.attach_vae): pass pre-encoded latents [B, 16, h, w]")
    pixels = pipeline.image_processor.preprocess(images).to(pipeline.device).to(pipeline.dtype)
    z = vae.encode(pixels).latent_dist.sample()
    return (z - vae.config.shift_factor) * vae.config.scaling_factor


def _ids_for(pipeline, latents: Tensor, n_tokens: int) -> Tensor:
    """Position ids of the packed tokens; the reference retries with the halved grid when its diffusers version counts
    the latent grid in pixels rather than in 2x2 patches (pipeline_tools.py:15-29)."""
    B, _, h, w = latents.shape
    for grid in ((h, w), (h // 2, w // 2)):
        ids = pipeline._prepare_latent_image_ids(B, grid[0], grid[1], pipeline.device, pipeline.dtype)
        if ids.shape[0] == n_tokens:
            break
    return ids


def encode_images(pipeline, images: Tensor):
    """pipeline_tools.py:7-30: VAE-encode -> (x - shift) * scale -> _pack_latents -> ids, as (tokens [B, N, 64], ids [N, 3]).

    With a VAE attached (`pipeline.attach_vae(...)`) pictures go through the native encoder; latents go
    straight to the native pack kernel."""
    latents = _to_latents(pipeline, images)
    tokens = pipeline._pack_latents(latents, *latents.shape)
    return tokens, _ids_for(pipeline, latents, tokens.shape[1])


def prepare_text_input(pipeline, prompts, max_sequence_length=512):
    """pipeline_tools.py:33-52: -> (prompt_embeds, pooled_prompt_embeds, text_ids); needs `pipeline.attach_text_encoders`."""
    return pipeline.encode_prompt(prompt=prompts, prompt_2=None, prompt_embeds=None, pooled_prompt_embeds=None,
                                  device=pipeline.device, num_images_per_prompt=1,
                                  max_sequence_length=max_sequence_length, lora_scale=None)
