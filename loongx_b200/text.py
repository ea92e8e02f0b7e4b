"""Native text encoders: what `FluxPipeline.encode_prompt` runs once per edit (generate.py:156-165,
pipeline_tools.py:33-52) — transformers' T5EncoderModel (T5 v1.1 XXL -> prompt_embeds [B, 512, 4096]) and CLIPTextModel
(CLIP-L -> pooled_prompt_embeds [B, 768]).  No attention mask reaches either model in diffusers 0.31.0's FluxPipeline.

Host side = weight packing and pointer plumbing: Linears on the tcgen05 GEMM (q|k|v and wi_0|wi_1 fused, residual adds in
the GEMM epilogue), the rest in csrc/text.cu.  The residual stream is bf16 (the reference's encoders run in the pipeline
dtype).  Tokenisation is host-side string work and is delegated to the tokenizer objects the caller supplies
(`transformers`' CLIPTokenizer / T5TokenizerFast when a checkpoint directory is given).  The oracle
(oracle/text_encoders.py, pinned against transformers) is never imported here; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L
from . import ops

_lib = L.lib
c_void_p, c_int32, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class SmallAttnDesc(C.Structure):
    _fields_ = [("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("ldq", c_int64), ("ldk", c_int64), ("ldv", c_int64),
                ("out", c_void_p), ("ldo", c_int64), ("bias", c_void_p), ("B", c_int32), ("H", c_int32), ("S", c_int32),
                ("head_dim", c_int32), ("causal", c_int32), ("scale", c_float), ("bias_relative", c_int32), ("reserved", c_int32)]


_lib.lx_embed_rows.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p]
_lib.lx_norm_rows.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_float, c_int32,
                              c_void_p]
_lib.lx_mul_rows.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p]
_lib.lx_attention_small.argtypes = [C.POINTER(SmallAttnDesc), c_void_p]


def _stream() -> int:
    return L.current_stream()


def _cuda(t: Optional[torch.Tensor]):
    if t is None:
        return None
    assert t.is_cuda, "loongx_b200.text needs CUDA tensors (there is no CPU fallback)"
    return t.data_ptr()


@dataclass
class T5Config:
    """google/t5-v1_1-xxl (FLUX.1-dev text_encoder_2/config.json)."""
    vocab_size: int = 32128
    d_model: int = 4096
    d_kv: int = 64
    num_heads: int = 64
    d_ff: int = 10240
    num_layers: int = 24
    num_buckets: int = 32
    max_distance: int = 128
    eps: float = 1e-6

    @staticmethod
    def from_json(cj: dict) -> "T5Config":
        if cj.get("feed_forward_proj", "gated-gelu") != "gated-gelu":
            raise NotImplementedError(f"feed_forward_proj={cj.get('feed_forward_proj')!r}: only gated-gelu (T5 v1.1) is built")
        return T5Config(cj["vocab_size"], cj["d_model"], cj["d_kv"], cj["num_heads"], cj["d_ff"], cj["num_layers"],
                        cj.get("relative_attention_num_buckets", 32), cj.get("relative_attention_max_distance", 128),
                        cj.get("layer_norm_epsilon", 1e-6))


@dataclass
class ClipTextConfig:
    """openai/clip-vit-large-patch14 text tower (FLUX.1-dev text_encoder/config.json)."""
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_layers: int = 12
    num_heads: int = 12
    max_positions: int = 77
    eps: float = 1e-5
    eos_token_id: int = 2

    @staticmethod
    def from_json(cj: dict) -> "ClipTextConfig":
        if cj.get("hidden_act", "quick_gelu") != "quick_gelu":
            raise NotImplementedError(f"hidden_act={cj.get('hidden_act')!r}: only quick_gelu (CLIP-L) is built")
        return ClipTextConfig(cj["vocab_size"], cj["hidden_size"], cj["intermediate_size"], cj["num_hidden_layers"],
                              cj["num_attention_heads"], cj.get("max_position_embeddings", 77), cj.get("layer_norm_eps", 1e-5),
                              cj.get("eos_token_id", 2))


def t5_relative_buckets(S: int, num_buckets: int, max_distance: int) -> torch.Tensor:
    """Bucket index of (query i, key j), bidirectional (integer bookkeeping; T5Attention._relative_position_bucket)."""
    rel = torch.arange(S)[None, :] - torch.arange(S)[:, None]
    nb = num_buckets // 2
    out = (rel > 0).long() * nb
    rel = rel.abs()
    exact = nb // 2
    far = exact + (torch.log(rel.float() / exact) / math.log(max_distance / exact) * (nb - exact)).long()
    return out + torch.where(rel < exact, rel, far.clamp(max=nb - 1))


def _bf(t, dev):
    return t.to(device=dev, dtype=torch.bfloat16).contiguous()


def _f32(t, dev):
    return t.to(device=dev, dtype=torch.float32).contiguous()


class _Blocks:
    """Shared execution helpers: activations are bf16 rows [B*S, D]."""

    def _init_common(self, device):
        self.device = torch.device(device)
        self._ones = torch.ones(16384, dtype=torch.bfloat16, device=self.device)
        self._meta = torch.zeros(64, 4, dtype=torch.int32, device=self.device)
        self.launches = 0

    def _tile_meta(self, M):
        need = (M + 127) // 128 + 2
        if self._meta.shape[0] < need:
            self._meta = torch.zeros(need, 4, dtype=torch.int32, device=self.device)
        return self._meta

    def _norm(self, x, gamma, beta, eps, rms):
        out = torch.empty_like(x)
        L.check(_lib.lx_norm_rows(_cuda(x), x.stride(0), _cuda(gamma), _cuda(beta), _cuda(out), out.stride(0), x.shape[0],
                                  x.shape[1], eps, int(rms), _stream()), "lx_norm_rows")
        self.launches += 1
        return out

    def _linear(self, x, w, bias, mode=L.EPI_BIAS, residual=None, out=None, **kw):
        if out is None:
            out = torch.empty(x.shape[0], w.shape[0], dtype=torch.bfloat16, device=self.device)
        if residual is not None:
            ops.gemm(x, w, bias, out, L.EPI_GATE_RESIDUAL, tile_meta=self._tile_meta(x.shape[0]), residual=residual,
                     gate=[self._ones[:w.shape[0]], None, None])
        else:
            ops.gemm(x, w, bias, out, mode, **kw)
        self.launches += 1
        return out

    def _attention(self, qkv, B, S, H, inner, bias, causal, scale, bias_relative=False):
        out = torch.empty(B * S, inner, dtype=torch.bfloat16, device=self.device)
        d = SmallAttnDesc()
        d.q, d.k, d.v = _cuda(qkv), _cuda(qkv[:, inner:]), _cuda(qkv[:, 2 * inner:])
        d.ldq = d.ldk = d.ldv = qkv.stride(0)
        d.out, d.ldo, d.bias = _cuda(out), out.stride(0), _cuda(bias)
        d.B, d.H, d.S, d.head_dim, d.causal, d.scale = B, H, S, inner // H, int(causal), scale
        d.bias_relative = int(bool(bias_relative))
        L.check(_lib.lx_attention_small(C.byref(d), _stream()), "lx_attention_small")
        self.launches += 1
        return out

    def _embed(self, table, ids, pos, period):
        n, D = ids.numel(), table.shape[1]
        ids32 = ids.to(device=self.device, dtype=torch.int32).contiguous().view(-1)
        out = torch.empty(n, D, dtype=torch.bfloat16, device=self.device)
        L.check(_lib.lx_embed_rows(_cuda(table), _cuda(ids32), _cuda(pos), period, _cuda(out), n, D, table.shape[0], _stream()),
                "lx_embed_rows")
        self.launches += 1
        return out


class NativeT5Encoder(_Blocks):
    """`pipeline.text_encoder_2`: input ids [B, S <= 512] -> last hidden state [B, S, d_model] (bf16)."""

    def __init__(self, cfg: T5Config, P: Dict[str, torch.Tensor], device="cuda"):
        if cfg.d_kv != 64:
            raise NotImplementedError(f"d_kv={cfg.d_kv}: the attention kernel is built for 64-wide heads")
        if cfg.d_ff % 256:
            raise NotImplementedError(f"d_ff={cfg.d_ff} must be a multiple of 256 (two-segment GEMM epilogue)")
        self._init_common(device)
        self.cfg = cfg
        dev = self.device
        emb = P["encoder.embed_tokens.weight"] if "encoder.embed_tokens.weight" in P else P["shared.weight"]
        self.embed = _bf(emb, dev)
        self.rel_bias = _f32(P["encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"], dev)  # [buckets, H]
        self.final_norm = _f32(P["encoder.final_layer_norm.weight"], dev)
        self.layers = []
        for i in range(cfg.num_layers):
            p = f"encoder.block.{i}.layer."
            a = p + "0.SelfAttention."
            self.layers.append(dict(
                ln1=_f32(P[p + "0.layer_norm.weight"], dev),
                qkv=_bf(torch.cat([P[a + "q.weight"], P[a + "k.weight"], P[a + "v.weight"]], 0), dev),
                o=_bf(P[a + "o.weight"], dev),
                ln2=_f32(P[p + "1.layer_norm.weight"], dev),
                wi=_bf(torch.cat([P[p + "1.DenseReluDense.wi_0.weight"], P[p + "1.DenseReluDense.wi_1.weight"]], 0), dev),
                wo=_bf(P[p + "1.DenseReluDense.wo.weight"], dev)))
        self._bias_cache: Dict[int, torch.Tensor] = {}

    def _bias(self, S: int) -> torch.Tensor:
        """Relative position bias as a distance table [H, 2S-1] (entry key - query + S - 1), shared by every layer (T5
        computes it in block 0 only)."""
        if S not in self._bias_cache:
            b = t5_relative_buckets(S, self.cfg.num_buckets, self.cfg.max_distance)
            by_distance = torch.cat([b[1:, 0].flip(0), b[0, :]]).to(self.device)  # distances -(S-1) .. S-1
            self._bias_cache[S] = self.rel_bias[by_distance].t().contiguous()
        return self._bias_cache[S]

    def __call__(self, input_ids: torch.Tensor, **_) -> Tuple[torch.Tensor]:
        cfg = self.cfg
        B, S = input_ids.shape
        inner, ff = cfg.num_heads * cfg.d_kv, cfg.d_ff
        h = self._embed(self.embed, input_ids, None, 0)
        bias = self._bias(S)
        for lw in self.layers:
            n = self._norm(h, lw["ln1"], None, cfg.eps, True)
            qkv = self._linear(n, lw["qkv"], None)
            a = self._attention(qkv, B, S, cfg.num_heads, inner, bias, False, 1.0, bias_relative=True)  # T5: no 1/sqrt(d) scaling
            h = self._linear(a, lw["o"], None, residual=h)
            n = self._norm(h, lw["ln2"], None, cfg.eps, True)
            g = torch.empty(B * S, 2 * ff, dtype=torch.bfloat16, device=self.device)
            # one launch: columns [0, ff) = gelu_new(wi_0 x), columns [ff, 2 ff) = wi_1 x
            self._linear(n, lw["wi"], None, mode=L.EPI_BIAS_GELU, out=g, n_split=ff, seg1=(L.EPI_BIAS, g, ff))
            L.check(_lib.lx_mul_rows(_cuda(g), g.stride(0), _cuda(g[:, ff:]), g.stride(0), _cuda(g), g.stride(0), B * S, ff,
                                     _stream()), "lx_mul_rows")
            self.launches += 1
            h = self._linear(g[:, :ff], lw["wo"], None, residual=h)
        out = self._norm(h, self.final_norm, None, cfg.eps, True)
        return (out.view(B, S, cfg.d_model),)


class NativeClipText(_Blocks):
    """`pipeline.text_encoder`: input ids [B, S <= 77] -> object with `.last_hidden_state` and `.pooler_output`."""

    class Output:
        def __init__(self, last_hidden_state, pooler_output):
            self.last_hidden_state, self.pooler_output = last_hidden_state, pooler_output

        def __getitem__(self, i):
            return (self.last_hidden_state, self.pooler_output)[i]

    def __init__(self, cfg: ClipTextConfig, P: Dict[str, torch.Tensor], device="cuda"):
        if cfg.hidden_size // cfg.num_heads != 64:
            raise NotImplementedError("the attention kernel is built for 64-wide heads")
        self._init_common(device)
        self.cfg = cfg
        dev, t = self.device, "text_model."
        self.tok = _bf(P[t + "embeddings.token_embedding.weight"], dev)
        self.pos = _bf(P[t + "embeddings.position_embedding.weight"], dev)
        self.final = (_f32(P[t + "final_layer_norm.weight"], dev), _f32(P[t + "final_layer_norm.bias"], dev))
        self.layers = []
        for i in range(cfg.num_layers):
            p = f"{t}encoder.layers.{i}."
            a = p + "self_attn."
            self.layers.append(dict(
                ln1=(_f32(P[p + "layer_norm1.weight"], dev), _f32(P[p + "layer_norm1.bias"], dev)),
                qkv=_bf(torch.cat([P[a + f"{n}_proj.weight"] for n in "qkv"], 0), dev),
                qkv_b=_f32(torch.cat([P[a + f"{n}_proj.bias"] for n in "qkv"], 0), dev),
                o=_bf(P[a + "out_proj.weight"], dev), o_b=_f32(P[a + "out_proj.bias"], dev),
                ln2=(_f32(P[p + "layer_norm2.weight"], dev), _f32(P[p + "layer_norm2.bias"], dev)),
                # quick_gelu(y) = y * sigmoid(1.702 y) = silu(1.702 y) / 1.702: the factor goes into fc1, its inverse into fc2
                fc1=_bf(P[p + "mlp.fc1.weight"].float() * 1.702, dev), fc1_b=_f32(P[p + "mlp.fc1.bias"].float() * 1.702, dev),
                fc2=_bf(P[p + "mlp.fc2.weight"].float() / 1.702, dev), fc2_b=_f32(P[p + "mlp.fc2.bias"], dev)))

    def __call__(self, input_ids: torch.Tensor, **_) -> "NativeClipText.Output":
        cfg = self.cfg
        B, S = input_ids.shape
        if S > cfg.max_positions:
            raise ValueError(f"sequence length {S} exceeds the {cfg.max_positions} learned positions")
        D, H = cfg.hidden_size, cfg.num_heads
        h = self._embed(self.tok, input_ids, self.pos, S)
        for lw in self.layers:
            n = self._norm(h, lw["ln1"][0], lw["ln1"][1], cfg.eps, False)
            qkv = self._linear(n, lw["qkv"], lw["qkv_b"])
            a = self._attention(qkv, B, S, H, D, None, True, (D // H) ** -0.5)
            h = self._linear(a, lw["o"], lw["o_b"], residual=h)
            n = self._norm(h, lw["ln2"][0], lw["ln2"][1], cfg.eps, False)
            m = self._linear(n, lw["fc1"], lw["fc1_b"], mode=L.EPI_BIAS_SILU)
            h = self._linear(m, lw["fc2"], lw["fc2_b"], residual=h)
        out = self._norm(h, self.final[0], self.final[1], cfg.eps, False).view(B, S, D)
        ids = input_ids.to(self.device)
        pos = ids.argmax(-1) if cfg.eos_token_id == 2 else (ids == cfg.eos_token_id).int().argmax(-1)
        return NativeClipText.Output(out, out[torch.arange(B, device=self.device), pos])


# ------------------------------------------------------------------------------------------------------------------
# checkpoint directories (FluxPipeline.from_pretrained layout)
# ------------------------------------------------------------------------------------------------------------------
def _read_safetensors_dir(d: str) -> Dict[str, torch.Tensor]:
    from safetensors import safe_open

    idx = os.path.join(d, "model.safetensors.index.json")
    if os.path.exists(idx):
        with open(idx) as f:
            files = sorted(set(json.load(f)["weight_map"].values()))
    else:
        files = ["model.safetensors"]
    P: Dict[str, torch.Tensor] = {}
    for fn in files:
        with safe_open(os.path.join(d, fn), framework="pt", device="cpu") as sf:
            for k in sf.keys():
                P[k] = sf.get_tensor(k)
    return P


def load_text_encoders(flux_path: str, device="cuda") -> Tuple[NativeClipText, NativeT5Encoder]:
    """<flux_path>/text_encoder (CLIP-L) and <flux_path>/text_encoder_2 (T5-XXL) -> native encoders."""
    with open(os.path.join(flux_path, "text_encoder", "config.json")) as f:
        ccfg = ClipTextConfig.from_json(json.load(f))
    with open(os.path.join(flux_path, "text_encoder_2", "config.json")) as f:
        tcfg = T5Config.from_json(json.load(f))
    clip = NativeClipText(ccfg, _read_safetensors_dir(os.path.join(flux_path, "text_encoder")), device)
    t5 = NativeT5Encoder(tcfg, _read_safetensors_dir(os.path.join(flux_path, "text_encoder_2")), device)
    return clip, t5


def load_tokenizers(flux_path: str):
    """CLIPTokenizer / T5TokenizerFast from <flux_path>/tokenizer, tokenizer_2 (host-side string processing: transformers)."""
    from transformers import AutoTokenizer

    return (AutoTokenizer.from_pretrained(os.path.join(flux_path, "tokenizer")),
            AutoTokenizer.from_pretrained(os.path.join(flux_path, "tokenizer_2")))


def tokenize(tokenizer, prompts: List[str], max_length: int) -> torch.Tensor:
    """diffusers' call: padding="max_length", truncation=True, return_tensors="pt" -> input ids [B, max_length]."""
    enc = tokenizer(prompts, padding="max_length", max_length=max_length, truncation=True, return_length=False,
                    return_overflowing_tokens=False, return_tensors="pt")
    return enc["input_ids"]
