"""Model geometry of the Flux MM-DiT that LoongX drives (FLUX.1-dev transformer/config.json; the reference reads it
through `flux_path`, train/config/seed_512.yaml:1) and the names/shapes of its Linear layers in diffusers
state-dict naming, so a diffusers-format checkpoint maps 1:1 onto the native weight container."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple


@dataclass
class FluxConfig:
    num_layers: int = 19
    num_single_layers: int = 38
    num_attention_heads: int = 24
    attention_head_dim: int = 128
    in_channels: int = 64
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    guidance_embeds: bool = True
    axes_dims_rope: Tuple[int, int, int] = (16, 56, 56)
    mlp_ratio: int = 4
    lora_rank: int = 4
    lora_alpha: float = 4.0

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    def validate(self) -> None:
        if self.attention_head_dim != 128:
            raise ValueError("the sm_100a kernels are specialised for attention_head_dim == 128")
        if self.num_attention_heads % 2 or self.inner_dim > 3072:
            raise ValueError("num_attention_heads must be even and inner_dim <= 3072")
        if sum(self.axes_dims_rope) != self.attention_head_dim:
            raise ValueError("axes_dims_rope must sum to the head dim")


# LoRA targets of train/config/seed_512.yaml:38 (the regex spelled out)
DOUBLE_LORA = ("norm1.linear", "attn.to_q", "attn.to_k", "attn.to_v", "attn.to_out.0", "ff.net.2")
SINGLE_LORA = ("norm.linear", "proj_mlp", "proj_out", "attn.to_q", "attn.to_k", "attn.to_v")


def lora_targets(cfg: FluxConfig) -> List[str]:
    names = ["x_embedder"]
    for i in range(cfg.num_layers):
        names += [f"transformer_blocks.{i}.{n}" for n in DOUBLE_LORA]
    for i in range(cfg.num_single_layers):
        names += [f"single_transformer_blocks.{i}.{n}" for n in SINGLE_LORA]
    return names


def linear_shapes(cfg: FluxConfig) -> Dict[str, Tuple[int, int]]:
    """Linear name -> (out_features, in_features)."""
    D, FF = cfg.inner_dim, cfg.inner_dim * cfg.mlp_ratio
    s: Dict[str, Tuple[int, int]] = {
        "x_embedder": (D, cfg.in_channels),
        "context_embedder": (D, cfg.joint_attention_dim),
        "time_text_embed.timestep_embedder.linear_1": (D, 256),
        "time_text_embed.timestep_embedder.linear_2": (D, D),
        "time_text_embed.text_embedder.linear_1": (D, cfg.pooled_projection_dim),
        "time_text_embed.text_embedder.linear_2": (D, D),
        "norm_out.linear": (2 * D, D),
        "proj_out": (cfg.in_channels, D),
    }
    if cfg.guidance_embeds:
        s["time_text_embed.guidance_embedder.linear_1"] = (D, 256)
        s["time_text_embed.guidance_embedder.linear_2"] = (D, D)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        s[p + "norm1.linear"] = (6 * D, D)
        s[p + "norm1_context.linear"] = (6 * D, D)
        for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
            s[p + "attn." + n] = (D, D)
        s[p + "ff.net.0.proj"] = (FF, D)
        s[p + "ff.net.2"] = (D, FF)
        s[p + "ff_context.net.0.proj"] = (FF, D)
        s[p + "ff_context.net.2"] = (D, FF)
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}."
        s[p + "norm.linear"] = (3 * D, D)
        s[p + "proj_mlp"] = (FF, D)
        s[p + "proj_out"] = (D, D + FF)
        for n in ("to_q", "to_k", "to_v"):
            s[p + "attn." + n] = (D, D)
    return s


def rmsnorm_names(cfg: FluxConfig) -> List[str]:
    names = []
    for i in range(cfg.num_layers):
        names += [f"transformer_blocks.{i}.attn.{n}.weight" for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k")]
    for i in range(cfg.num_single_layers):
        names += [f"single_transformer_blocks.{i}.attn.{n}.weight" for n in ("norm_q", "norm_k")]
    return names
