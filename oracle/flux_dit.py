"""ORACLE — test infrastructure only (tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import this; the product path under loongx_b200/ and src/ never does).

Pure-PyTorch restatement of the LoongX / OminiControl DiT forward for any device / dtype:

  * reference functions restated 1:1 in structure (same concat order, same fp32 islands, same LoRA masking):
      attn_forward            /root/reference/src/flux/block.py:7-176
      block_forward           /root/reference/src/flux/block.py:179-278
      single_block_forward    /root/reference/src/flux/block.py:281-339
      tranformer_forward      /root/reference/src/flux/transformer.py:47-252
      enable_lora semantics   /root/reference/src/flux/lora_controller.py:5-43
  * third-party arithmetic the reference only *calls* (absent from /root/reference; diffusers pinned ==0.31.0 in
    train/requirements.txt:1, peft unpinned): restated from the published diffusers 0.31.0 / peft algorithms as
    recorded in SURVEY.md App. A.1-A.8 (AdaLayerNormZero/-Single/-Continuous, RMSNorm, FluxPosEmbed,
    apply_rotary_emb, Timesteps/TimestepEmbedding/PixArtAlphaTextProjection, FeedForward(gelu-approximate),
    peft LoRA Linear).

PINNING.  (1) Pinned to the reference's OWN source: oracle/ref_harness.py imports block.py / transformer.py /
lora_controller.py / generate.py from /root/reference and executes them on the CPU; this restatement reproduces their
outputs bit-for-bit on 7 model_config / c_factor / c_t / no-condition variants and 4 generate() runs
(tests/golden/ref_v1.npz written by tests/golden/make_ref_golden.py, checked by tests/test_reference_pins_cpu.py).
(2) The third-party arithmetic: diffusers / peft are not installed and the reference ships no golden vectors or
checkpoints, so in (1) the diffusers modules the reference receives as arguments are stand-ins written
from the same published algorithm (App. A).  Round 2 PINNED that arithmetic to an executable third-party implementation
of the same network: Black Forest Labs' FLUX model as shipped in the image's `torchtitan` package (EmbedND RoPE,
timestep_embedding, MLPEmbedder, Modulation, QK-RMSNorm, Double / SingleStreamBlock, LastLayer).
tests/golden/make_dit_bfl_golden.py maps this file's diffusers-named parameters onto that model (the published
diffusers <-> BFL key correspondence, incl. the shift / scale swap of norm_out) and writes tests/golden/dit_bfl_v1.npz;
tests/test_oracle_cpu.py replays it and re-runs the model live: `tranformer_forward` agrees to 3e-7 relL2 both without a
condition branch (the stock forward) and through the reference's three-stream forward with the condition tokens at
c_t = t and LoRA B = 0 (where they are arithmetically more image tokens).  Left PARITY UNPINNED: the guidance embedder
(torchtitan's copy has none; it is the timestep MLPEmbedder once more, covered structurally), peft's LoRA Linear (closed
form y = W x + s B A x, checked in tests/test_oracle_cpu.py) and s4torch (oracle/cs3_dgf.py).

Parameters live in a flat dict keyed by the diffusers state-dict names of SURVEY.md App. A.9; LoRA factors are stored
as "<linear>.lora_A.weight" [r, in] and "<linear>.lora_B.weight" [out, r] with scaling = lora_alpha / r.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


@dataclass
class FluxConfig:
    """FLUX.1-dev transformer/config.json — tiny variants are used by the tests."""

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


# LoRA target regex of train/config/seed_512.yaml:38, spelled out
DOUBLE_LORA = ("norm1.linear", "attn.to_q", "attn.to_k", "attn.to_v", "attn.to_out.0", "ff.net.2")
SINGLE_LORA = ("norm.linear", "proj_mlp", "proj_out", "attn.to_q", "attn.to_k", "attn.to_v")


def lora_target_names(cfg: FluxConfig):
    names = ["x_embedder"]
    for i in range(cfg.num_layers):
        names += [f"transformer_blocks.{i}.{n}" for n in DOUBLE_LORA]
    for i in range(cfg.num_single_layers):
        names += [f"single_transformer_blocks.{i}.{n}" for n in SINGLE_LORA]
    return names


def linear_shapes(cfg: FluxConfig) -> Dict[str, Tuple[int, int]]:
    """name -> (out_features, in_features) of every nn.Linear of FluxTransformer2DModel (App. A.3/A.5/A.9)."""
    D, FF = cfg.inner_dim, cfg.inner_dim * cfg.mlp_ratio
    s = {
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


def rmsnorm_names(cfg: FluxConfig):
    names = []
    for i in range(cfg.num_layers):
        names += [f"transformer_blocks.{i}.attn.{n}.weight" for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k")]
    for i in range(cfg.num_single_layers):
        names += [f"single_transformer_blocks.{i}.attn.{n}.weight" for n in ("norm_q", "norm_k")]
    return names


def init_params(cfg: FluxConfig, seed: int = 1234, dtype=torch.float32, device="cpu", w_std: float = 0.02,
                bias_std: float = 0.0, lora_b_std: float = 0.02, alias_blocks: bool = False) -> Params:
    """Synthetic weights of SURVEY.md §8d: Linear.weight ~ N(0, w_std^2), RMSNorm weight 1 (+ small jitter so the
    weight is exercised), LoRA A ~ N(0, 1/r) (peft 'gaussian'), LoRA B ~ N(0, lora_b_std^2) (non-zero so the
    condition branch is exercised).  alias_blocks=True reuses block 0's tensors for all blocks (CPU baseline on
    hosts that cannot hold 47.6 GB of fp32 weights; identical FLOPs)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    P: Params = {}

    def randn(*shape, std=1.0):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype).to(device)

    shapes = linear_shapes(cfg)
    targets = set(lora_target_names(cfg))

    def make_linear(name):
        o, i = shapes[name]
        P[name + ".weight"] = randn(o, i, std=w_std)
        P[name + ".bias"] = randn(o, std=bias_std) if bias_std > 0 else torch.zeros(o, dtype=dtype, device=device)
        if name in targets and cfg.lora_rank > 0:
            P[name + ".lora_A.weight"] = randn(cfg.lora_rank, i, std=1.0 / cfg.lora_rank)
            P[name + ".lora_B.weight"] = randn(o, cfg.lora_rank, std=lora_b_std)

    def alias(dst_prefix, src_prefix):
        for k in [k for k in P if k.startswith(src_prefix)]:
            P[dst_prefix + k[len(src_prefix):]] = P[k]

    for name in shapes:
        if alias_blocks:
            parts = name.split(".")
            if parts[0] in ("transformer_blocks", "single_transformer_blocks") and parts[1] != "0":
                continue
        make_linear(name)
    for n in rmsnorm_names(cfg):
        if alias_blocks and n.split(".")[1] != "0":
            continue
        P[n] = (1.0 + 0.1 * torch.randn(cfg.attention_head_dim, generator=g)).to(dtype).to(device)
    if alias_blocks:
        for i in range(1, cfg.num_layers):
            alias(f"transformer_blocks.{i}.", "transformer_blocks.0.")
        for i in range(1, cfg.num_single_layers):
            alias(f"single_transformer_blocks.{i}.", "single_transformer_blocks.0.")
    return P


# ------------------------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------------------------
def linear(P: Params, name: str, x: torch.Tensor, lora: bool, cfg: FluxConfig) -> torch.Tensor:
    """nn.Linear, wrapped by a peft LoRA layer when targeted: y = base(x) + B(A(x)) * scaling (App. A.8).
    `lora=False` is the state inside `enable_lora(..., activated=False)` (scaling multiplied by 0,
    lora_controller.py:22-28)."""
    y = F.linear(x, P[name + ".weight"], P[name + ".bias"])
    a = P.get(name + ".lora_A.weight")
    if lora and a is not None:
        scaling = cfg.lora_alpha / cfg.lora_rank
        y = y + F.linear(F.linear(x, a), P[name + ".lora_B.weight"]) * scaling
    return y


def layer_norm(x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), None, None, 1e-6)


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """diffusers RMSNorm (App. A.1): variance in fp32, cast to the weight dtype if half/bf16, then scale."""
    in_dtype = x.dtype
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    x = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        x = x.to(weight.dtype)
    x = x * weight
    return x if weight.dtype in (torch.float16, torch.bfloat16) else x.to(in_dtype)


def ada_layer_norm_zero(P, name, x, emb, lora, cfg):
    """AdaLayerNormZero.forward (App. A.2) -> (x_mod, gate_msa, shift_mlp, scale_mlp, gate_mlp)."""
    e = linear(P, name + ".linear", F.silu(emb), lora, cfg)
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = e.chunk(6, dim=1)
    x = layer_norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
    return x, gate_msa, shift_mlp, scale_mlp, gate_mlp


def ada_layer_norm_zero_single(P, name, x, emb, lora, cfg):
    e = linear(P, name + ".linear", F.silu(emb), lora, cfg)
    shift, scale, gate = e.chunk(3, dim=1)
    return layer_norm(x) * (1 + scale[:, None]) + shift[:, None], gate


def ada_layer_norm_continuous(P, name, x, emb, cfg):
    """norm_out: scale FIRST, then shift (App. A.2)."""
    e = linear(P, name + ".linear", F.silu(emb).to(x.dtype), False, cfg)
    scale, shift = e.chunk(2, dim=1)
    return layer_norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]


def rope_tables(ids: torch.Tensor, axes_dim=(16, 56, 56), theta: float = 10000.0):
    """FluxPosEmbed.forward (App. A.4): ids [S,3] -> (cos, sin) [S, sum(axes_dim)] fp32, float64 internally."""
    pos = ids.float()
    cos_out, sin_out = [], []
    for i, d in enumerate(axes_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64, device=ids.device) / d))
        ang = torch.outer(pos[:, i].to(torch.float64), freqs)
        cos_out.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin_out.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos_out, dim=-1), torch.cat(sin_out, dim=-1)


def apply_rotary_emb(x: torch.Tensor, freqs) -> torch.Tensor:
    """diffusers apply_rotary_emb, use_real=True, unbind_dim=-1 (App. A.4): interleaved pairs, fp32 math."""
    cos, sin = freqs
    cos, sin = cos[None, None].to(x.device), sin[None, None].to(x.device)
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    compute = torch.float64 if x.dtype == torch.float64 else torch.float32
    return (x.to(compute) * cos.to(compute) + x_rot.to(compute) * sin.to(compute)).to(x.dtype)


def timestep_sinusoid(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0) (App. A.5)."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    e = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(e), torch.sin(e)], dim=-1)


def time_text_embed(P, cfg, timestep, guidance, pooled):
    """CombinedTimestep(Guidance)TextProjEmbeddings (App. A.5)."""
    pre = "time_text_embed."

    def mlp(name, x):
        h = F.silu(linear(P, pre + name + ".linear_1", x, False, cfg))
        return linear(P, pre + name + ".linear_2", h, False, cfg)

    t_emb = mlp("timestep_embedder", timestep_sinusoid(timestep).to(pooled.dtype))
    if guidance is not None:
        g_emb = mlp("guidance_embedder", timestep_sinusoid(guidance).to(pooled.dtype))
        cond = t_emb + g_emb
    else:
        cond = t_emb
    return cond + mlp("text_embedder", pooled)


# ------------------------------------------------------------------------------------------------------------------
# block.py restated
# ------------------------------------------------------------------------------------------------------------------
def attn_forward(P, cfg, prefix, hidden_states, encoder_hidden_states=None, condition_latents=None,
                 image_rotary_emb=None, cond_rotary_emb=None, model_config: Optional[dict] = None,
                 c_factor: Optional[float] = None):
    """block.py:7-176.  `prefix` names the diffusers Attention module ('transformer_blocks.3.attn')."""
    model_config = model_config or {}
    H, hd = cfg.num_attention_heads, cfg.attention_head_dim
    latent_lora = model_config.get("latent_lora", False)
    B = hidden_states.shape[0]

    def heads(x):
        return x.view(B, -1, H, hd).transpose(1, 2)

    query = heads(linear(P, prefix + ".to_q", hidden_states, latent_lora, cfg))  # :23-36
    key = heads(linear(P, prefix + ".to_k", hidden_states, latent_lora, cfg))
    value = heads(linear(P, prefix + ".to_v", hidden_states, latent_lora, cfg))
    query = rms_norm(query, P[prefix + ".norm_q.weight"])  # :38-41
    key = rms_norm(key, P[prefix + ".norm_k.weight"])

    if encoder_hidden_states is not None:  # :44-72
        eq = heads(linear(P, prefix + ".add_q_proj", encoder_hidden_states, False, cfg))
        ek = heads(linear(P, prefix + ".add_k_proj", encoder_hidden_states, False, cfg))
        ev = heads(linear(P, prefix + ".add_v_proj", encoder_hidden_states, False, cfg))
        eq = rms_norm(eq, P[prefix + ".norm_added_q.weight"])
        ek = rms_norm(ek, P[prefix + ".norm_added_k.weight"])
        query = torch.cat([eq, query], dim=2)
        key = torch.cat([ek, key], dim=2)
        value = torch.cat([ev, value], dim=2)

    if image_rotary_emb is not None:  # :74-78
        query = apply_rotary_emb(query, image_rotary_emb)
        key = apply_rotary_emb(key, image_rotary_emb)

    if condition_latents is not None:  # :80-104 (LoRA active)
        cq = heads(linear(P, prefix + ".to_q", condition_latents, True, cfg))
        ck = heads(linear(P, prefix + ".to_k", condition_latents, True, cfg))
        cv = heads(linear(P, prefix + ".to_v", condition_latents, True, cfg))
        cq = rms_norm(cq, P[prefix + ".norm_q.weight"])
        ck = rms_norm(ck, P[prefix + ".norm_k.weight"])
        if cond_rotary_emb is not None:
            cq = apply_rotary_emb(cq, cond_rotary_emb)
            ck = apply_rotary_emb(ck, cond_rotary_emb)
        query = torch.cat([query, cq], dim=2)
        key = torch.cat([key, ck], dim=2)
        value = torch.cat([value, cv], dim=2)

    attention_mask = None  # :106-128
    if condition_latents is not None:
        n_c = condition_latents.shape[1]
        S = query.shape[2]
        if not model_config.get("union_cond_attn", True):
            attention_mask = torch.ones(S, S, device=query.device, dtype=torch.bool)
            attention_mask[-n_c:, :-n_c] = False
            attention_mask[:-n_c, -n_c:] = False
        elif model_config.get("independent_condition", False):
            attention_mask = torch.ones(S, S, device=query.device, dtype=torch.bool)
            attention_mask[-n_c:, :-n_c] = False
        if c_factor is not None:
            attention_mask = torch.zeros(S, S, device=query.device, dtype=query.dtype)
            bias = math.log(c_factor)
            attention_mask[-n_c:, :-n_c] = bias
            attention_mask[:-n_c, -n_c:] = bias

    out = F.scaled_dot_product_attention(query, key, value, dropout_p=0.0, is_causal=False, attn_mask=attention_mask)
    out = out.transpose(1, 2).reshape(B, -1, H * hd).to(query.dtype)  # :132-135

    if encoder_hidden_states is not None:  # :137-167
        n_t = encoder_hidden_states.shape[1]
        if condition_latents is not None:
            n_c = condition_latents.shape[1]
            enc, hid, cond = out[:, :n_t], out[:, n_t:-n_c], out[:, -n_c:]
        else:
            enc, hid, cond = out[:, :n_t], out[:, n_t:], None
        hid = linear(P, prefix + ".to_out.0", hid, latent_lora, cfg)
        enc = linear(P, prefix + ".to_add_out", enc, False, cfg)
        if cond is not None:
            cond = linear(P, prefix + ".to_out.0", cond, True, cfg)
            return hid, enc, cond
        return hid, enc
    elif condition_latents is not None:  # :168-174
        n_c = condition_latents.shape[1]
        return out[:, :-n_c], out[:, -n_c:]
    return out


def feed_forward(P, cfg, prefix, x, lora_down):
    """FeedForward(activation_fn='gelu-approximate') (App. A.3): net.0.proj -> GELU(tanh) -> net.2."""
    h = F.gelu(linear(P, prefix + ".net.0.proj", x, False, cfg), approximate="tanh")
    return linear(P, prefix + ".net.2", h, lora_down, cfg)


def block_forward(P, cfg, idx, hidden_states, encoder_hidden_states, condition_latents, temb, cond_temb,
                  cond_rotary_emb=None, image_rotary_emb=None, model_config: Optional[dict] = None, c_factor=None):
    """block.py:179-278 -> (encoder_hidden_states, hidden_states, condition_latents)."""
    model_config = model_config or {}
    pre = f"transformer_blocks.{idx}"
    latent_lora = model_config.get("latent_lora", False)
    use_cond = condition_latents is not None
    norm_h, gate_msa, shift_mlp, scale_mlp, gate_mlp = ada_layer_norm_zero(P, pre + ".norm1", hidden_states, temb,
                                                                           latent_lora, cfg)
    norm_e, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = ada_layer_norm_zero(
        P, pre + ".norm1_context", encoder_hidden_states, temb, False, cfg)
    if use_cond:
        norm_c, cond_gate_msa, cond_shift_mlp, cond_scale_mlp, cond_gate_mlp = ada_layer_norm_zero(
            P, pre + ".norm1", condition_latents, cond_temb, True, cfg)

    result = attn_forward(P, cfg, pre + ".attn", norm_h, norm_e, norm_c if use_cond else None,
                          image_rotary_emb=image_rotary_emb, cond_rotary_emb=cond_rotary_emb if use_cond else None,
                          model_config=model_config, c_factor=c_factor)
    attn_output, context_attn_output = result[:2]
    cond_attn_output = result[2] if use_cond else None

    hidden_states = hidden_states + gate_msa.unsqueeze(1) * attn_output  # :224-225
    encoder_hidden_states = encoder_hidden_states + c_gate_msa.unsqueeze(1) * context_attn_output
    if use_cond:
        cond_attn_output = cond_gate_msa.unsqueeze(1) * cond_attn_output
        condition_latents = condition_latents + cond_attn_output
        if model_config.get("add_cond_attn", False):
            hidden_states = hidden_states + cond_attn_output

    norm_h = layer_norm(hidden_states) * (1 + scale_mlp[:, None]) + shift_mlp[:, None]  # :238-253
    norm_e = layer_norm(encoder_hidden_states) * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
    if use_cond:
        norm_c = layer_norm(condition_latents) * (1 + cond_scale_mlp[:, None]) + cond_shift_mlp[:, None]

    ff_output = gate_mlp.unsqueeze(1) * feed_forward(P, cfg, pre + ".ff", norm_h, latent_lora)  # :256-266
    context_ff_output = c_gate_mlp.unsqueeze(1) * feed_forward(P, cfg, pre + ".ff_context", norm_e, False)
    if use_cond:
        cond_ff_output = cond_gate_mlp.unsqueeze(1) * feed_forward(P, cfg, pre + ".ff", norm_c, True)

    hidden_states = hidden_states + ff_output
    encoder_hidden_states = encoder_hidden_states + context_ff_output
    if use_cond:
        condition_latents = condition_latents + cond_ff_output
    if encoder_hidden_states.dtype == torch.float16:
        encoder_hidden_states = encoder_hidden_states.clip(-65504, 65504)
    return encoder_hidden_states, hidden_states, condition_latents if use_cond else None


def single_block_forward(P, cfg, idx, hidden_states, temb, image_rotary_emb=None, condition_latents=None,
                         cond_temb=None, cond_rotary_emb=None, model_config: Optional[dict] = None, c_factor=None):
    """block.py:281-339."""
    model_config = model_config or {}
    pre = f"single_transformer_blocks.{idx}"
    latent_lora = model_config.get("latent_lora", False)
    using_cond = condition_latents is not None
    residual = hidden_states
    norm_h, gate = ada_layer_norm_zero_single(P, pre + ".norm", hidden_states, temb, latent_lora, cfg)
    mlp_h = F.gelu(linear(P, pre + ".proj_mlp", norm_h, latent_lora, cfg), approximate="tanh")
    if using_cond:
        residual_cond = condition_latents
        norm_c, cond_gate = ada_layer_norm_zero_single(P, pre + ".norm", condition_latents, cond_temb, True, cfg)
        mlp_c = F.gelu(linear(P, pre + ".proj_mlp", norm_c, True, cfg), approximate="tanh")

    attn_output = attn_forward(P, cfg, pre + ".attn", norm_h, None, norm_c if using_cond else None,
                               image_rotary_emb=image_rotary_emb,
                               cond_rotary_emb=cond_rotary_emb if using_cond else None, model_config=model_config,
                               c_factor=c_factor)
    if using_cond:
        attn_output, cond_attn_output = attn_output

    hidden_states = torch.cat([attn_output, mlp_h], dim=2)  # :325-329
    hidden_states = residual + gate.unsqueeze(1) * linear(P, pre + ".proj_out", hidden_states, latent_lora, cfg)
    if using_cond:
        condition_latents = torch.cat([cond_attn_output, mlp_c], dim=2)
        condition_latents = residual_cond + cond_gate.unsqueeze(1) * linear(P, pre + ".proj_out", condition_latents,
                                                                            True, cfg)
    if hidden_states.dtype == torch.float16:
        hidden_states = hidden_states.clip(-65504, 65504)
    return hidden_states if not using_cond else (hidden_states, condition_latents)


def tranformer_forward(P, cfg, condition_latents, condition_ids, condition_type_ids=None,
                       model_config: Optional[dict] = None, c_t=0, *, hidden_states, encoder_hidden_states,
                       pooled_projections, timestep, img_ids, txt_ids, guidance=None, c_factor=None,
                       controlnet_block_samples=None, controlnet_single_block_samples=None):
    """transformer.py:47-252 -> noise prediction [B, N_img, in_channels].  condition_type_ids is ignored exactly like
    the reference (transformer.py:133 is commented out).  controlnet_*_samples: lists of [B, N_img, inner_dim] residuals
    added to the image stream after the blocks (transformer.py:172-181, 230-239)."""
    model_config = model_config or {}
    use_condition = condition_latents is not None
    latent_lora = model_config.get("latent_lora", False)

    hidden_states = linear(P, "x_embedder", hidden_states, latent_lora, cfg)  # :91-93
    condition_latents = linear(P, "x_embedder", condition_latents, True, cfg) if use_condition else None

    timestep = timestep.to(hidden_states.dtype) * 1000  # :95-100
    guidance = guidance.to(hidden_states.dtype) * 1000 if guidance is not None else None
    temb = time_text_embed(P, cfg, timestep, guidance, pooled_projections)  # :102-106
    cond_temb = time_text_embed(P, cfg, torch.ones_like(timestep) * c_t * 1000, guidance, pooled_projections)
    encoder_hidden_states = linear(P, "context_embedder", encoder_hidden_states, False, cfg)  # :115

    if txt_ids.ndim == 3:
        txt_ids = txt_ids[0]
    if img_ids.ndim == 3:
        img_ids = img_ids[0]
    ids = torch.cat((txt_ids, img_ids), dim=0)  # :130-134
    image_rotary_emb = rope_tables(ids, cfg.axes_dims_rope)
    cond_rotary_emb = rope_tables(condition_ids, cfg.axes_dims_rope) if use_condition else None

    for i in range(cfg.num_layers):  # :138-170
        encoder_hidden_states, hidden_states, condition_latents = block_forward(
            P, cfg, i, hidden_states, encoder_hidden_states, condition_latents if use_condition else None, temb,
            cond_temb if use_condition else None, cond_rotary_emb, image_rotary_emb, model_config, c_factor)
        if controlnet_block_samples is not None:  # :172-181
            interval_control = int(math.ceil(cfg.num_layers / len(controlnet_block_samples)))
            hidden_states = hidden_states + controlnet_block_samples[i // interval_control]
    n_txt = encoder_hidden_states.shape[1]
    hidden_states = torch.cat([encoder_hidden_states, hidden_states], dim=1)  # :182
    for i in range(cfg.num_single_layers):  # :184-228
        result = single_block_forward(P, cfg, i, hidden_states, temb, image_rotary_emb,
                                      condition_latents if use_condition else None,
                                      cond_temb if use_condition else None, cond_rotary_emb, model_config, c_factor)
        if use_condition:
            hidden_states, condition_latents = result
        else:
            hidden_states = result
        if controlnet_single_block_samples is not None:  # :230-239 (image rows of the joint [txt | img] stream)
            interval_control = int(math.ceil(cfg.num_single_layers / len(controlnet_single_block_samples)))
            hidden_states = torch.cat([hidden_states[:, :n_txt],
                                       hidden_states[:, n_txt:] + controlnet_single_block_samples[i // interval_control]], dim=1)
    hidden_states = hidden_states[:, n_txt:, ...]  # :241
    hidden_states = ada_layer_norm_continuous(P, "norm_out", hidden_states, temb, cfg)  # :243
    return linear(P, "proj_out", hidden_states, False, cfg)  # :244
