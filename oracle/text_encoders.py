"""ORACLE (test infrastructure only) — CPU restatement of the two text encoders behind `FluxPipeline.encode_prompt`
(SURVEY.md §8f.4).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.

Reference call sites: generate.py:156-165 `self.encode_prompt(prompt=..., prompt_2=..., max_sequence_length=512)` and
pipeline_tools.py:33-52 `prepare_text_input`; both land in diffusers 0.31.0's FluxPipeline:
  _get_clip_prompt_embeds   CLIPTokenizer(padding="max_length", max_length=77) -> CLIPTextModel(ids).pooler_output   [B, 768]
  _get_t5_prompt_embeds     T5TokenizerFast(padding="max_length", max_length=512) -> T5EncoderModel(ids)[0]          [B, 512, 4096]
(no attention mask is passed to either model: padding tokens are attended like any other token).

The arithmetic lives in a THIRD-PARTY package, `transformers` (unpinned in the reference's requirements.txt).  Unlike
diffusers it IS importable in this image (transformers 5.5), so this restatement is PINNED: tests/test_text_cpu.py builds
`transformers.T5EncoderModel` / `CLIPTextModel` with seeded random weights, copies their state dicts into the functions
below and requires agreement to fp32 rounding.  State-dict naming is transformers' own.
  T5 v1.1 (google/t5-v1_1-xxl: d_model 4096, d_kv 64, 64 heads, d_ff 10240, 24 layers, gated-gelu, 32 relative-position
  buckets, max distance 128, eps 1e-6): modeling_t5.py T5LayerNorm / T5Attention (no 1/sqrt(d) scaling, shared relative
  position bias of layer 0) / T5DenseGatedActDense (gelu_new) / T5Stack.
  CLIP text (openai/clip-vit-large-patch14: 768 wide, 12 layers, 12 heads, 77 positions, quick_gelu, eps 1e-5):
  modeling_clip.py CLIPTextEmbeddings / CLIPEncoderLayer (pre-LN, causal mask) / CLIPTextTransformer pooling.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F


@dataclass
class T5Cfg:
    vocab_size: int = 32128
    d_model: int = 4096
    d_kv: int = 64
    num_heads: int = 64
    d_ff: int = 10240
    num_layers: int = 24
    num_buckets: int = 32
    max_distance: int = 128
    eps: float = 1e-6


@dataclass
class ClipCfg:
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_layers: int = 12
    num_heads: int = 12
    max_positions: int = 77
    eps: float = 1e-5
    eos_token_id: int = 2  # the legacy config value: pooled token = argmax(input_ids)


# ---------------------------------------------------------------------------------------------------------------------
# T5 encoder
# ---------------------------------------------------------------------------------------------------------------------
def t5_relative_buckets(S: int, num_buckets: int = 32, max_distance: int = 128) -> torch.Tensor:
    """T5Attention._relative_position_bucket(bidirectional=True) for all (query i, key j): [S, S] int64."""
    ctx = torch.arange(S)[:, None]
    mem = torch.arange(S)[None, :]
    rel = mem - ctx
    nb = num_buckets // 2
    buckets = (rel > 0).long() * nb
    rel = rel.abs()
    max_exact = nb // 2
    is_small = rel < max_exact
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(is_small, rel, large)


def t5_layer_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    var = x.float().pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def gelu_new(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def t5_encode(P: Dict[str, torch.Tensor], ids: torch.Tensor, cfg: T5Cfg) -> torch.Tensor:
    """input ids [B, S] -> last hidden state [B, S, d_model] (T5EncoderModel(ids)[0], no attention mask)."""
    B, S = ids.shape
    H, dk = cfg.num_heads, cfg.d_kv
    h = P["encoder.embed_tokens.weight"][ids]
    bias = P["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"][
        t5_relative_buckets(S, cfg.num_buckets, cfg.max_distance)]  # [S, S, H]
    bias = bias.permute(2, 0, 1)[None]  # [1, H, S, S]
    for i in range(cfg.num_layers):
        p = f"encoder.block.{i}.layer."
        n = t5_layer_norm(h, P[p + "0.layer_norm.weight"], cfg.eps)
        q = F.linear(n, P[p + "0.SelfAttention.q.weight"]).view(B, S, H, dk).transpose(1, 2)
        k = F.linear(n, P[p + "0.SelfAttention.k.weight"]).view(B, S, H, dk).transpose(1, 2)
        v = F.linear(n, P[p + "0.SelfAttention.v.weight"]).view(B, S, H, dk).transpose(1, 2)
        a = torch.softmax((q @ k.transpose(2, 3) + bias).float(), dim=-1).to(v.dtype) @ v
        h = h + F.linear(a.transpose(1, 2).reshape(B, S, H * dk), P[p + "0.SelfAttention.o.weight"])
        n = t5_layer_norm(h, P[p + "1.layer_norm.weight"], cfg.eps)
        g = gelu_new(F.linear(n, P[p + "1.DenseReluDense.wi_0.weight"])) * F.linear(n, P[p + "1.DenseReluDense.wi_1.weight"])
        h = h + F.linear(g, P[p + "1.DenseReluDense.wo.weight"])
    return t5_layer_norm(h, P["encoder.final_layer_norm.weight"], cfg.eps)


def t5_init(cfg: T5Cfg, seed: int = 1234) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, std: torch.randn(*s, generator=g) * std  # noqa: E731
    inner = cfg.num_heads * cfg.d_kv
    P = {"encoder.embed_tokens.weight": r(cfg.vocab_size, cfg.d_model, std=1.0),
         "encoder.final_layer_norm.weight": 1.0 + r(cfg.d_model, std=0.1)}
    P["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"] = r(cfg.num_buckets, cfg.num_heads, std=0.5)
    for i in range(cfg.num_layers):
        p = f"encoder.block.{i}.layer."
        P[p + "0.SelfAttention.q.weight"] = r(inner, cfg.d_model, std=(cfg.d_model * cfg.d_kv) ** -0.5)
        P[p + "0.SelfAttention.k.weight"] = r(inner, cfg.d_model, std=cfg.d_model ** -0.5)
        P[p + "0.SelfAttention.v.weight"] = r(inner, cfg.d_model, std=cfg.d_model ** -0.5)
        P[p + "0.SelfAttention.o.weight"] = r(cfg.d_model, inner, std=inner ** -0.5)
        P[p + "0.layer_norm.weight"] = 1.0 + r(cfg.d_model, std=0.1)
        P[p + "1.DenseReluDense.wi_0.weight"] = r(cfg.d_ff, cfg.d_model, std=cfg.d_model ** -0.5)
        P[p + "1.DenseReluDense.wi_1.weight"] = r(cfg.d_ff, cfg.d_model, std=cfg.d_model ** -0.5)
        P[p + "1.DenseReluDense.wo.weight"] = r(cfg.d_model, cfg.d_ff, std=cfg.d_ff ** -0.5)
        P[p + "1.layer_norm.weight"] = 1.0 + r(cfg.d_model, std=0.1)
    return P


# ---------------------------------------------------------------------------------------------------------------------
# CLIP text model
# ---------------------------------------------------------------------------------------------------------------------
def clip_encode(P: Dict[str, torch.Tensor], ids: torch.Tensor, cfg: ClipCfg) -> Tuple[torch.Tensor, torch.Tensor]:
    """input ids [B, S<=77] -> (last_hidden_state [B, S, D], pooler_output [B, D])."""
    B, S = ids.shape
    H, D = cfg.num_heads, cfg.hidden_size
    dh = D // H
    t = "text_model."
    h = P[t + "embeddings.token_embedding.weight"][ids] + P[t + "embeddings.position_embedding.weight"][:S]
    causal = torch.full((S, S), float("-inf")).triu(1)
    for i in range(cfg.num_layers):
        p = f"{t}encoder.layers.{i}."
        n = F.layer_norm(h, (D,), P[p + "layer_norm1.weight"], P[p + "layer_norm1.bias"], cfg.eps)
        q = (F.linear(n, P[p + "self_attn.q_proj.weight"], P[p + "self_attn.q_proj.bias"]) * dh ** -0.5).view(B, S, H, dh).transpose(1, 2)
        k = F.linear(n, P[p + "self_attn.k_proj.weight"], P[p + "self_attn.k_proj.bias"]).view(B, S, H, dh).transpose(1, 2)
        v = F.linear(n, P[p + "self_attn.v_proj.weight"], P[p + "self_attn.v_proj.bias"]).view(B, S, H, dh).transpose(1, 2)
        a = torch.softmax(q @ k.transpose(2, 3) + causal, dim=-1) @ v
        h = h + F.linear(a.transpose(1, 2).reshape(B, S, D), P[p + "self_attn.out_proj.weight"], P[p + "self_attn.out_proj.bias"])
        n = F.layer_norm(h, (D,), P[p + "layer_norm2.weight"], P[p + "layer_norm2.bias"], cfg.eps)
        m = F.linear(n, P[p + "mlp.fc1.weight"], P[p + "mlp.fc1.bias"])
        m = m * torch.sigmoid(1.702 * m)  # quick_gelu
        h = h + F.linear(m, P[p + "mlp.fc2.weight"], P[p + "mlp.fc2.bias"])
    h = F.layer_norm(h, (D,), P[t + "final_layer_norm.weight"], P[t + "final_layer_norm.bias"], cfg.eps)
    if cfg.eos_token_id == 2:  # legacy configs (openai/clip-vit-large-patch14): the EOS token has the largest id
        pos = ids.argmax(-1)
    else:
        pos = (ids == cfg.eos_token_id).int().argmax(-1)
    return h, h[torch.arange(B), pos]


def clip_init(cfg: ClipCfg, seed: int = 1234) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, std: torch.randn(*s, generator=g) * std  # noqa: E731
    D, Fd, t = cfg.hidden_size, cfg.intermediate_size, "text_model."
    P = {t + "embeddings.token_embedding.weight": r(cfg.vocab_size, D, std=0.5),
         t + "embeddings.position_embedding.weight": r(cfg.max_positions, D, std=0.5),
         t + "final_layer_norm.weight": 1.0 + r(D, std=0.1), t + "final_layer_norm.bias": r(D, std=0.1)}
    for i in range(cfg.num_layers):
        p = f"{t}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            P[p + f"self_attn.{n}.weight"] = r(D, D, std=D ** -0.5)
            P[p + f"self_attn.{n}.bias"] = r(D, std=0.05)
        for n in ("layer_norm1", "layer_norm2"):
            P[p + n + ".weight"] = 1.0 + r(D, std=0.1)
            P[p + n + ".bias"] = r(D, std=0.1)
        P[p + "mlp.fc1.weight"], P[p + "mlp.fc1.bias"] = r(Fd, D, std=D ** -0.5), r(Fd, std=0.05)
        P[p + "mlp.fc2.weight"], P[p + "mlp.fc2.bias"] = r(D, Fd, std=Fd ** -0.5), r(D, std=0.05)
    return P
