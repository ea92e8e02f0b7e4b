"""Host side of the native DiT engine: weight packing (diffusers-named tensors -> fused bf16 panels), plan / workspace
allocation through torch, and thin calls into lx_dit_prepare / lx_dit_step.

No arithmetic of the hot path happens here — only one-time layout work at load (concatenating q/k/v(/proj_mlp)
weights, merging W + (alpha/r) B A for the LoRA-active row group, stacking the AdaLN linears) and pointer plumbing per
call.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L
from . import ops
from .config import FluxConfig, linear_shapes, lora_targets, rmsnorm_names

c_void_p, c_int32, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float


# ---------------------------------------------------------------------------------------------------------------
# ctypes mirrors of include/loongx_b200.h
# ---------------------------------------------------------------------------------------------------------------
class LxLinear(C.Structure):
    _fields_ = [("w", c_void_p), ("ldw", c_int64), ("bias", c_void_p), ("w_lora", c_void_p),
                ("n", c_int32), ("k", c_int32)]


class LxDoubleBlock(C.Structure):
    _fields_ = [("qkv", LxLinear), ("qkv_ctx", LxLinear), ("out", LxLinear), ("out_ctx", LxLinear),
                ("ff_up", LxLinear), ("ff_down", LxLinear), ("ff_ctx_up", LxLinear), ("ff_ctx_down", LxLinear),
                ("norm_q", c_void_p), ("norm_k", c_void_p), ("norm_added_q", c_void_p), ("norm_added_k", c_void_p)]


class LxSingleBlock(C.Structure):
    _fields_ = [("qkv_mlp", LxLinear), ("proj_out", LxLinear), ("norm_q", c_void_p), ("norm_k", c_void_p)]


class LxDitModel(C.Structure):
    _fields_ = [("num_layers", c_int32), ("num_single_layers", c_int32), ("heads", c_int32), ("in_channels", c_int32),
                ("joint_dim", c_int32), ("pooled_dim", c_int32), ("guidance_embeds", c_int32), ("reserved", c_int32),
                ("axes_dim", c_int32 * 3), ("reserved2", c_int32),
                ("x_embedder", LxLinear), ("context_embedder", LxLinear),
                ("time_1", LxLinear), ("time_2", LxLinear), ("guid_1", LxLinear), ("guid_2", LxLinear),
                ("text_1", LxLinear), ("text_2", LxLinear),
                ("mod_img", LxLinear), ("mod_txt", LxLinear), ("mod_single", LxLinear),
                ("norm_out", LxLinear), ("proj_out", LxLinear),
                ("double_blocks", c_void_p), ("single_blocks", c_void_p)]


class LxDitPlan(C.Structure):
    _fields_ = [("B", c_int32), ("n_txt", c_int32), ("n_img", c_int32), ("n_cond", c_int32),
                ("T", c_int32), ("mask_mode", c_int32), ("latent_lora", c_int32), ("add_cond_attn", c_int32),
                ("cross_bias", c_float), ("reserved", c_int32),
                ("tile_meta", c_void_p), ("out_row_base", c_void_p), ("rope", c_void_p),
                ("X", c_void_p), ("XN", c_void_p), ("Q", c_void_p), ("K", c_void_p), ("V", c_void_p),
                ("scratch", c_void_p), ("X0_txt", c_void_p), ("X0_cond", c_void_p),
                ("emb_tmp", c_void_p), ("sin_tmp", c_void_p), ("silu_t", c_void_p), ("silu_c", c_void_p),
                ("mod_img", c_void_p), ("mod_txt", c_void_p), ("mod_single", c_void_p), ("mod_out", c_void_p),
                ("mod_cond_img", c_void_p), ("mod_cond_single", c_void_p),
                ("t_dev", c_void_p), ("g_dev", c_void_p), ("pad", c_int32 * 3), ("cond_cached", c_int32),
                ("kv_block_stride", c_int64)]


_lib = L.lib
_lib.lx_dit_prepare.argtypes = [C.POINTER(LxDitModel), C.POINTER(LxDitPlan), c_void_p, c_void_p, c_void_p,
                                C.POINTER(c_float), C.POINTER(c_float), c_float, c_void_p]
_lib.lx_dit_embed.argtypes = [C.POINTER(LxDitModel), C.POINTER(LxDitPlan), c_void_p, c_void_p]
_lib.lx_dit_step.argtypes = [C.POINTER(LxDitModel), C.POINTER(LxDitPlan), c_int32, c_void_p, c_void_p, c_void_p]
_lib.lx_dit_head.argtypes = [C.POINTER(LxDitModel), C.POINTER(LxDitPlan), c_int32, c_void_p, c_void_p]
_lib.lx_add_rows.argtypes = [c_void_p, C.c_int64, c_void_p, C.c_int64, c_int32, c_int32, c_void_p]
_lib.lx_dit_double_block.argtypes = [C.POINTER(LxDitModel), C.POINTER(LxDitPlan), c_int32, c_int32, c_void_p]
_lib.lx_dit_single_block.argtypes = [C.POINTER(LxDitModel), C.POINTER(LxDitPlan), c_int32, c_int32, c_void_p]
_lib.lx_euler_step.argtypes = [c_void_p, c_void_p, c_void_p, c_float, c_int64, c_void_p]
_lib.lx_rope_table.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, C.c_double, c_void_p]
_lib.lx_pack_latents.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]


def _stream() -> int:
    return L.current_stream()


# ---------------------------------------------------------------------------------------------------------------
# weights
# ---------------------------------------------------------------------------------------------------------------
def random_params(cfg: FluxConfig, device, seed: int = 1234, w_std: float = 0.02, bias_std: float = 0.0,
                  lora_b_std: float = 0.02, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Synthetic weights with the reference architecture (there are no checkpoints on the box): Linear ~ N(0, w_std^2),
    RMSNorm weight 1 + 0.1 N(0,1), LoRA A ~ N(0, 1/r) ('gaussian' init), LoRA B ~ N(0, lora_b_std^2).  Generated
    directly on `device` so the 12 B-parameter model never exists on the host."""
    g = torch.Generator(device=device).manual_seed(seed)
    P: Dict[str, torch.Tensor] = {}
    targets = set(lora_targets(cfg))

    def randn(*shape, std):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * std).to(dtype)

    for name, (o, i) in linear_shapes(cfg).items():
        P[name + ".weight"] = randn(o, i, std=w_std)
        P[name + ".bias"] = randn(o, std=bias_std) if bias_std > 0 else torch.zeros(o, device=device, dtype=dtype)
        if name in targets and cfg.lora_rank > 0:
            P[name + ".lora_A.weight"] = randn(cfg.lora_rank, i, std=1.0 / cfg.lora_rank)
            P[name + ".lora_B.weight"] = randn(o, cfg.lora_rank, std=lora_b_std)
    for n in rmsnorm_names(cfg):
        P[n] = (1.0 + 0.1 * torch.randn(cfg.attention_head_dim, generator=g, device=device)).to(dtype)
    return P


class PackedLinear:
    """One (possibly fused / stacked) Linear in the native layout; keeps the tensors alive.

    `lora` lists the LoRA-targeted sub-Linears of the panel as (name, first_row, rows, A fp32 [r, K], B fp32 [rows, r]);
    the training step (loongx_b200/train.py) differentiates w.r.t. those factors, re-merges `w_lora` after an optimizer
    step and multiplies by the transposed panels `wT` / `w_loraT` in the dX GEMMs."""

    def __init__(self, w: torch.Tensor, bias: Optional[torch.Tensor], w_lora: Optional[torch.Tensor], lora=None,
                 scaling: float = 1.0):
        self.w, self.bias, self.w_lora = w, bias, w_lora
        self.lora = lora or []
        self.scaling = scaling
        self.wT: Optional[torch.Tensor] = None
        self.w_loraT: Optional[torch.Tensor] = None

    def c(self) -> LxLinear:
        s = LxLinear()
        s.w, s.ldw = self.w.data_ptr(), self.w.stride(0)
        s.bias = self.bias.data_ptr() if self.bias is not None else None
        s.w_lora = self.w_lora.data_ptr() if self.w_lora is not None else None
        s.n, s.k = self.w.shape
        return s


def pack_linear(P: Dict[str, torch.Tensor], names: Sequence[str], cfg: FluxConfig, device, *,
                pop: bool = False) -> PackedLinear:
    """Stack the Linear layers `names` along the output dimension (q|k|v(|proj_mlp), all blocks' AdaLN linears, ...).
    If any of them carries LoRA factors, also build the merged panel W + (alpha/r) lora_B lora_A (fp32 sum, one bf16
    rounding) that the LoRA-active row group of the GEMM reads (peft LoRA Linear)."""
    get = (lambda key: P.pop(key)) if pop else (lambda key: P[key])
    has_lora = any((n + ".lora_A.weight") in P for n in names)
    has_bias = (names[0] + ".bias") in P
    ws, merged, factors = [], [], []
    row0, scaling = 0, 1.0
    for n in names:
        lora = (n + ".lora_A.weight") in P
        w = get(n + ".weight").to(device=device, dtype=torch.bfloat16)
        ws.append(w)
        if has_lora:
            if lora:
                a = get(n + ".lora_A.weight").to(device=device, dtype=torch.float32).contiguous()
                b = get(n + ".lora_B.weight").to(device=device, dtype=torch.float32).contiguous()
                scaling = cfg.lora_alpha / a.shape[0]
                merged.append(torch.addmm(w.float(), b, a, alpha=scaling).to(torch.bfloat16))
                factors.append((n, row0, w.shape[0], a, b))
            else:
                merged.append(w)
        row0 += w.shape[0]
    bias = torch.cat([get(n + ".bias").to(device=device, dtype=torch.float32) for n in names]) if has_bias else None
    w = torch.cat(ws, 0).contiguous() if len(ws) > 1 else ws[0].contiguous()
    w_lora = None
    if has_lora:
        w_lora = torch.cat(merged, 0).contiguous() if len(merged) > 1 else merged[0].contiguous()
    return PackedLinear(w, bias, w_lora, factors, scaling)


class DitWeights:
    """Native weight container: packed device tensors + the ctypes model struct handed to the C ABI."""

    def __init__(self, P: Dict[str, torch.Tensor], cfg: FluxConfig, device="cuda", consume: bool = False):
        cfg.validate()
        self.cfg = cfg
        self.device = torch.device(device)
        dev = self.device
        def pk(names, **kw):
            self._last_names = list(names)
            return pack_linear(P, names, cfg, dev, pop=consume, **kw)

        f32 = lambda key: (P.pop(key) if consume else P[key]).to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        self.keep: List[object] = []
        self.named: Dict[str, PackedLinear] = {}  # "x_embedder", "double.3.qkv", "single.7.proj_out", ... -> panel
        self.layout: Dict[str, List[str]] = {}    # panel / norm key -> diffusers module names stacked in it (for export)
        m = LxDitModel()
        m.num_layers, m.num_single_layers = cfg.num_layers, cfg.num_single_layers
        m.heads, m.in_channels = cfg.num_attention_heads, cfg.in_channels
        m.joint_dim, m.pooled_dim = cfg.joint_attention_dim, cfg.pooled_projection_dim
        m.guidance_embeds = int(cfg.guidance_embeds)
        for i in range(3):
            m.axes_dim[i] = cfg.axes_dims_rope[i]

        def put(field: str, pl: PackedLinear):
            self.keep.append(pl)
            self.named[field] = pl
            self.layout[field] = self._last_names
            setattr(m, field, pl.c())

        put("x_embedder", pk(["x_embedder"]))
        put("context_embedder", pk(["context_embedder"]))
        put("time_1", pk(["time_text_embed.timestep_embedder.linear_1"]))
        put("time_2", pk(["time_text_embed.timestep_embedder.linear_2"]))
        if cfg.guidance_embeds:
            put("guid_1", pk(["time_text_embed.guidance_embedder.linear_1"]))
            put("guid_2", pk(["time_text_embed.guidance_embedder.linear_2"]))
        put("text_1", pk(["time_text_embed.text_embedder.linear_1"]))
        put("text_2", pk(["time_text_embed.text_embedder.linear_2"]))
        put("mod_img", pk([f"transformer_blocks.{i}.norm1.linear" for i in range(cfg.num_layers)]))
        put("mod_txt", pk([f"transformer_blocks.{i}.norm1_context.linear" for i in range(cfg.num_layers)]))
        put("mod_single", pk([f"single_transformer_blocks.{i}.norm.linear" for i in range(cfg.num_single_layers)]))
        put("norm_out", pk(["norm_out.linear"]))
        put("proj_out", pk(["proj_out"]))

        self.dbl = (LxDoubleBlock * max(cfg.num_layers, 1))()
        for i in range(cfg.num_layers):
            p = f"transformer_blocks.{i}."
            blk = self.dbl[i]
            for field, names in (
                ("qkv", [p + "attn.to_q", p + "attn.to_k", p + "attn.to_v"]),
                ("qkv_ctx", [p + "attn.add_q_proj", p + "attn.add_k_proj", p + "attn.add_v_proj"]),
                ("out", [p + "attn.to_out.0"]), ("out_ctx", [p + "attn.to_add_out"]),
                ("ff_up", [p + "ff.net.0.proj"]), ("ff_down", [p + "ff.net.2"]),
                ("ff_ctx_up", [p + "ff_context.net.0.proj"]), ("ff_ctx_down", [p + "ff_context.net.2"]),
            ):
                pl = pk(names)
                self.keep.append(pl)
                self.named[f"double.{i}.{field}"] = pl
                self.layout[f"double.{i}.{field}"] = list(names)
                setattr(blk, field, pl.c())
            for field in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
                t = f32(p + f"attn.{field}.weight")
                self.keep.append(t)
                self.named[f"double.{i}.{field}"] = t
                self.layout[f"double.{i}.{field}"] = [p + f"attn.{field}.weight"]
                setattr(blk, field, t.data_ptr())
        self.sgl = (LxSingleBlock * max(cfg.num_single_layers, 1))()
        for i in range(cfg.num_single_layers):
            p = f"single_transformer_blocks.{i}."
            blk = self.sgl[i]
            pl = pk([p + "attn.to_q", p + "attn.to_k", p + "attn.to_v", p + "proj_mlp"])
            self.keep.append(pl)
            self.named[f"single.{i}.qkv_mlp"] = pl
            self.layout[f"single.{i}.qkv_mlp"] = [p + "attn.to_q", p + "attn.to_k", p + "attn.to_v", p + "proj_mlp"]
            blk.qkv_mlp = pl.c()
            pl = pk([p + "proj_out"])
            self.keep.append(pl)
            self.named[f"single.{i}.proj_out"] = pl
            self.layout[f"single.{i}.proj_out"] = [p + "proj_out"]
            blk.proj_out = pl.c()
            for field in ("norm_q", "norm_k"):
                t = f32(p + f"attn.{field}.weight")
                self.keep.append(t)
                self.named[f"single.{i}.{field}"] = t
                self.layout[f"single.{i}.{field}"] = [p + f"attn.{field}.weight"]
                setattr(blk, field, t.data_ptr())
        m.double_blocks = C.cast(self.dbl, c_void_p)
        m.single_blocks = C.cast(self.sgl, c_void_p)
        self.model = m

    def export_params(self, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
        """Inverse of the packing: flat dict in diffusers naming (`<module>.weight/.bias`, RMSNorm weights, LoRA factors as
        `<module>.lora_A.weight` / `.lora_B.weight`), views / small copies of the packed panels."""
        shapes = linear_shapes(self.cfg)
        out: Dict[str, torch.Tensor] = {}
        for key, names in self.layout.items():
            obj = self.named[key]
            if not isinstance(obj, PackedLinear):
                out[names[0]] = obj.to(dtype)
                continue
            r0 = 0
            for n in names:
                rows = shapes[n][0]
                out[n + ".weight"] = obj.w[r0:r0 + rows].to(dtype)
                if obj.bias is not None:
                    out[n + ".bias"] = obj.bias[r0:r0 + rows].to(dtype)
                r0 += rows
            for (n, _, _, A, Bw) in obj.lora:
                out[n + ".lora_A.weight"], out[n + ".lora_B.weight"] = A, Bw
        return out

    def param_bytes(self) -> int:
        tot = 0
        for k in self.keep:
            if isinstance(k, PackedLinear):
                tot += k.w.numel() * 2 + (k.w_lora.numel() * 2 if k.w_lora is not None else 0)
        return tot


# ---------------------------------------------------------------------------------------------------------------
# plan (geometry + workspace)
# ---------------------------------------------------------------------------------------------------------------
MASK_NONE, MASK_NO_UNION, MASK_INDEPENDENT = 0, 1, 2


def mask_mode_from_config(model_config: Optional[dict]) -> int:
    """block.py:106-120: union_cond_attn=False wins over independent_condition."""
    model_config = model_config or {}
    if not model_config.get("union_cond_attn", True):
        return MASK_NO_UNION
    if model_config.get("independent_condition", False):
        return MASK_INDEPENDENT
    return MASK_NONE


def _pad128(n: int) -> int:
    return (n + 127) // 128 * 128


class DitPlan:
    """Geometry + workspace of one batch of edits.  Stream lengths are arbitrary (the reference accepts any H, W divisible
    by 16 and any prompt length); internally every stream is padded to a multiple of 128 tokens — the row-tile size of
    every kernel — and the padding keys are masked inside the attention kernels.  `nt, ni, nc` are the caller's lengths,
    `ntp, nip, ncp` the padded ones; inputs are zero-padded on the way in and the padding rows dropped on the way out
    (layout plumbing)."""

    def __init__(self, weights: DitWeights, B: int, n_txt: int, n_img: int, n_cond: int, T: int = 1,
                 model_config: Optional[dict] = None, c_factor: Optional[float] = None, cache_cond: bool = False):
        """cache_cond=True: under model_config.independent_condition the condition stream does not depend
        on the denoise step, so prepare() runs it once, keeps every block's condition keys / values, and step() processes
        the text + image rows only (-40 % rows at 512x512).  Ignored when the configuration does not allow it."""
        cfg = weights.cfg
        dev = weights.device
        if n_txt <= 0 or n_img <= 0 or n_cond < 0:
            raise ValueError(f"bad stream lengths {(n_txt, n_img, n_cond)}")
        model_config = model_config or {}
        self.weights, self.B, self.nt, self.ni, self.nc, self.T = weights, B, n_txt, n_img, n_cond, T
        self.ntp, self.nip, self.ncp = _pad128(n_txt), _pad128(n_img), _pad128(n_cond)
        self.padded = (self.ntp, self.nip, self.ncp) != (n_txt, n_img, n_cond)
        n_txt, n_img, n_cond = self.ntp, self.nip, self.ncp  # everything below is in padded tokens
        D, H = cfg.inner_dim, cfg.num_attention_heads
        S = n_txt + n_img + n_cond
        R = B * S
        L_, Ls = cfg.num_layers, cfg.num_single_layers
        bf = dict(device=dev, dtype=torch.bfloat16)
        M = T * B + B
        self.cache_cond = bool(cache_cond and n_cond > 0 and c_factor is None and
                               mask_mode_from_config(model_config) == MASK_INDEPENDENT and
                               not model_config.get("add_cond_attn", False))
        kv_slots = (L_ + Ls) if self.cache_cond else 1  # one K / V buffer per block when the condition part is kept
        self.buf = dict(
            tile_meta=ops.make_tile_meta(B, n_txt, n_img, n_cond, dev),
            out_row_base=ops.make_out_row_base(B, n_txt, n_img, n_cond, dev),
            rope=torch.zeros((S, 64, 2), device=dev, dtype=torch.float32),
            X=torch.zeros((R, D), **bf),
            XN=torch.zeros((R, D), **bf),
            Q=torch.zeros((B, H, S, 128), **bf), K=torch.zeros((kv_slots, B, H, S, 128), **bf)[0],
            V=torch.zeros((kv_slots, B, H, S, 128), **bf)[0],  # views of slot 0; slot b = base + b*kv_block_stride
            scratch=torch.zeros((R, 5 * D), **bf),
            X0_txt=torch.zeros((B * n_txt, D), **bf),
            X0_cond=torch.zeros((max(B * n_cond, 1), D), **bf),
            emb_tmp=torch.zeros((4, M, D), **bf),
            sin_tmp=torch.zeros((M, 256), **bf),
            silu_t=torch.zeros((T * B, D), **bf),
            silu_c=torch.zeros((B, D), **bf),
            mod_img=torch.zeros((T * B, max(L_, 1) * 6 * D), **bf),
            mod_txt=torch.zeros((T * B, max(L_, 1) * 6 * D), **bf),
            mod_single=torch.zeros((T * B, max(Ls, 1) * 3 * D), **bf),
            mod_out=torch.zeros((T * B, 2 * D), **bf),
            mod_cond_img=torch.zeros((B, max(L_, 1) * 6 * D), **bf),
            mod_cond_single=torch.zeros((B, max(Ls, 1) * 3 * D), **bf),
            t_dev=torch.zeros((M,), device=dev, dtype=torch.float32),
            g_dev=torch.zeros((M,), device=dev, dtype=torch.float32),
        )
        p = LxDitPlan()
        p.B, p.n_txt, p.n_img, p.n_cond, p.T = B, n_txt, n_img, n_cond, T
        p.mask_mode = mask_mode_from_config(model_config)
        p.latent_lora = int(bool(model_config.get("latent_lora", False)))
        p.add_cond_attn = int(bool(model_config.get("add_cond_attn", False)))
        p.cross_bias = math.log(c_factor) if c_factor is not None else 0.0
        for k, v in self.buf.items():
            setattr(p, k, v.data_ptr())
        p.pad[0], p.pad[1], p.pad[2] = self.ntp - self.nt, self.nip - self.ni, self.ncp - self.nc
        p.cond_cached = 0
        p.kv_block_stride = B * H * S * 128 if self.cache_cond else 0
        self.plan = p
        self.has_rope = False
        if self.padded:  # staging for the zero-padded latents / predictions of step()
            self._lat_pad = torch.zeros((B, self.nip, cfg.in_channels), **bf)
            self._out_pad = torch.zeros((B, self.nip, cfg.in_channels), **bf)

    @staticmethod
    def _pad_tokens(x: torch.Tensor, n_pad: int) -> torch.Tensor:
        """[B, n, C] -> [B, n_pad, C] with zero rows appended (no-op when already that long)."""
        if x.shape[1] == n_pad:
            return x
        out = torch.zeros((x.shape[0], n_pad, x.shape[2]), device=x.device, dtype=x.dtype)
        out[:, :x.shape[1]].copy_(x)
        return out

    def set_ids(self, txt_ids: torch.Tensor, img_ids: torch.Tensor, cond_ids: Optional[torch.Tensor]) -> None:
        """RoPE table of the joint [txt | img | cond] sequence (transformer.py:130-134); ids are step-invariant so
        this runs once per edit instead of once per forward."""
        cfg = self.weights.cfg
        parts = [(txt_ids, self.nt, self.ntp), (img_ids, self.ni, self.nip)] + \
            ([(cond_ids, self.nc, self.ncp)] if cond_ids is not None else [])
        padded = []
        for x, n, n_pad in parts:
            x = x.to(device=self.weights.device, dtype=torch.float32)
            assert x.shape == (n, 3), (tuple(x.shape), n)
            padded.append(x)
            if n_pad > n:
                padded.append(torch.zeros((n_pad - n, 3), device=x.device))  # padding tokens: position 0, masked as keys
        ids = torch.cat(padded, 0).contiguous()
        assert ids.shape == (self.ntp + self.nip + self.ncp, 3), ids.shape
        a = cfg.axes_dims_rope
        L.check(_lib.lx_rope_table(ids.data_ptr(), self.buf["rope"].data_ptr(), ids.shape[0], a[0], a[1], a[2], 10000.0,
                                   _stream()), "lx_rope_table")
        self._ids_keepalive = ids
        self.has_rope = True

    # -- native calls ------------------------------------------------------------------------------------------
    def prepare(self, prompt_embeds: torch.Tensor, pooled: torch.Tensor, cond_latents: Optional[torch.Tensor],
                timesteps: Sequence[float], guidance: Optional[Sequence[float]], c_t: float = 0.0) -> None:
        """timesteps: T*B values in (0,1] ordered step-major; guidance: B values or None."""
        w = self.weights
        assert self.has_rope, "call set_ids() first"
        assert len(timesteps) == self.T * self.B
        pe = prompt_embeds.to(torch.bfloat16).contiguous()
        po = pooled.to(torch.bfloat16).contiguous()
        cl = cond_latents.to(torch.bfloat16).contiguous() if cond_latents is not None else None
        assert pe.shape == (self.B, self.nt, w.cfg.joint_attention_dim), pe.shape
        assert cl is None or cl.shape == (self.B, self.nc, w.cfg.in_channels), cl.shape
        pe = self._pad_tokens(pe, self.ntp)
        cl = self._pad_tokens(cl, self.ncp) if cl is not None else None
        assert po.shape == (self.B, w.cfg.pooled_projection_dim)
        ts = (c_float * len(timesteps))(*[float(t) for t in timesteps])
        gs = (c_float * self.B)(*[float(x) for x in guidance]) if guidance is not None else None
        L.check(_lib.lx_dit_prepare(C.byref(w.model), C.byref(self.plan), pe.data_ptr(), po.data_ptr(),
                                    cl.data_ptr() if cl is not None else None, ts, gs, float(c_t), _stream()),
                "lx_dit_prepare")
        self._prep_keepalive = (pe, po, cl)
        if self.cache_cond:
            # one full pass with every block writing its own K / V buffer: the condition rows of each buffer are now the
            # block's step-invariant condition keys / values (the text / image rows are rewritten by every step)
            self.plan.cond_cached = 0
            dummy = torch.zeros((self.B, self.nip if self.padded else self.ni, w.cfg.in_channels), device=w.device,
                                dtype=torch.bfloat16)
            L.check(_lib.lx_dit_embed(C.byref(w.model), C.byref(self.plan), dummy.data_ptr(), _stream()), "lx_dit_embed")
            for i in range(w.cfg.num_layers):
                self.double_block(0, i)
            for i in range(w.cfg.num_single_layers):
                self.single_block(0, i)
            self.plan.cond_cached = 1

    def step(self, step: int, latents: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        w = self.weights
        assert latents.dtype == torch.bfloat16 and latents.is_contiguous()
        assert latents.shape == (self.B, self.ni, w.cfg.in_channels), latents.shape
        if out is None:
            out = torch.empty_like(latents)
        if self.padded:
            self._lat_pad[:, :self.ni].copy_(latents)
            src, dst = self._lat_pad, self._out_pad
        else:
            src, dst = latents, out
        L.check(_lib.lx_dit_step(C.byref(w.model), C.byref(self.plan), step, src.data_ptr(), dst.data_ptr(),
                                 _stream()), "lx_dit_step")
        if self.padded:
            out.copy_(self._out_pad[:, :self.ni])
        return out

    def step_with_residuals(self, step: int, latents: torch.Tensor, block_samples: Optional[Sequence[torch.Tensor]] = None,
                            single_block_samples: Optional[Sequence[torch.Tensor]] = None,
                            out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """step() with controlnet residuals (transformer.py:172-181, 230-239): after double block i the image stream gets
        `block_samples[i // ceil(n_blocks / len(block_samples))]` added, after single block i the image rows of the joint
        stream get `single_block_samples[...]`; samples are [B, n_img, inner_dim].  The forward runs block by block through
        the reference-granularity entry points (lx_dit_embed / lx_dit_double_block / lx_dit_single_block / lx_dit_head) with
        lx_add_rows between them."""
        w = self.weights
        cfg = w.cfg
        assert latents.dtype == torch.bfloat16 and latents.is_contiguous()
        assert latents.shape == (self.B, self.ni, cfg.in_channels), latents.shape
        if out is None:
            out = torch.empty_like(latents)
        D = cfg.inner_dim
        img = self.split_streams()[1]  # [B, n_img, D] view of the image rows of plan->X

        def add(samples, i, n_blocks):
            if samples is None:
                return
            interval = -(-n_blocks // len(samples))  # int(np.ceil(len(blocks) / len(samples)))
            r = samples[i // interval].to(device=w.device, dtype=torch.bfloat16)
            assert r.shape == (self.B, self.ni, D), (tuple(r.shape), (self.B, self.ni, D))
            for b in range(self.B):  # (a padded plan leaves a gap between the batches' image rows)
                rb = r[b] if r[b].stride(1) == 1 and r[b].stride(0) % 8 == 0 else r[b].contiguous()
                L.check(_lib.lx_add_rows(img[b].data_ptr(), img[b].stride(0), rb.data_ptr(), rb.stride(0), self.ni, D,
                                         _stream()), "lx_add_rows")

        self.embed(latents)
        for i in range(cfg.num_layers):
            self.double_block(step, i)
            add(block_samples, i, cfg.num_layers)
        for i in range(cfg.num_single_layers):
            self.single_block(step, i)
            add(single_block_samples, i, cfg.num_single_layers)
        dst = self._out_pad if self.padded else out
        L.check(_lib.lx_dit_head(C.byref(w.model), C.byref(self.plan), step, dst.data_ptr(), _stream()), "lx_dit_head")
        if self.padded:
            out.copy_(self._out_pad[:, :self.ni])
        return out

    def embed(self, latents: torch.Tensor) -> None:
        if self.padded:
            self._lat_pad[:, :self.ni].copy_(latents)
            latents = self._lat_pad
        L.check(_lib.lx_dit_embed(C.byref(self.weights.model), C.byref(self.plan), latents.data_ptr(), _stream()),
                "lx_dit_embed")

    def double_block(self, step: int, block: int) -> None:
        L.check(_lib.lx_dit_double_block(C.byref(self.weights.model), C.byref(self.plan), step, block, _stream()),
                "lx_dit_double_block")

    def single_block(self, step: int, block: int) -> None:
        L.check(_lib.lx_dit_single_block(C.byref(self.weights.model), C.byref(self.plan), step, block, _stream()),
                "lx_dit_single_block")

    # -- stream-major row layout <-> reference [B, N, D] tensors (test / shim helpers, pure views + copies) ---------
    def split_streams(self):
        X, B, D = self.buf["X"], self.B, self.weights.cfg.inner_dim
        rt, ri = B * self.ntp, B * self.nip
        txt = X[:rt].view(B, self.ntp, D)[:, :self.nt]
        img = X[rt:rt + ri].view(B, self.nip, D)[:, :self.ni]
        cond = X[rt + ri:].view(B, self.ncp, D)[:, :self.nc] if self.nc else None
        return txt, img, cond


def euler_step(latents: torch.Tensor, noise_pred: torch.Tensor, dt: float, out: Optional[torch.Tensor] = None):
    """scheduler.step (generate.py:349) on bf16 tensors: out = bf16(float(latents) + dt * float(noise_pred))."""
    assert latents.dtype == torch.bfloat16 and noise_pred.dtype == torch.bfloat16
    assert latents.is_contiguous() and noise_pred.is_contiguous()
    if out is None:
        out = torch.empty_like(latents)
    L.check(_lib.lx_euler_step(latents.data_ptr(), noise_pred.data_ptr(), out.data_ptr(), float(dt), latents.numel(),
                               _stream()), "lx_euler_step")
    return out


def pack_latents(x: torch.Tensor) -> torch.Tensor:
    """FluxPipeline._pack_latents: [B, C, h, w] -> [B, (h/2)(w/2), 4C]."""
    assert x.is_cuda and x.is_contiguous() and x.element_size() in (2, 4)
    B, Cc, h, w = x.shape
    out = torch.empty((B, (h // 2) * (w // 2), Cc * 4), device=x.device, dtype=x.dtype)
    L.check(_lib.lx_pack_latents(x.data_ptr(), out.data_ptr(), B, Cc, h, w, x.element_size(), 0, _stream()),
            "lx_pack_latents")
    return out


def unpack_latents(x: torch.Tensor, height: int, width: int, vae_scale_factor: int = 16) -> torch.Tensor:
    """FluxPipeline._unpack_latents (diffusers 0.31: vae_scale_factor = 16): [B, N, 4C] -> [B, C, H/8, W/8]."""
    assert x.is_cuda and x.is_contiguous() and x.element_size() in (2, 4)
    B, N, ch = x.shape
    h, w = 2 * (height // vae_scale_factor), 2 * (width // vae_scale_factor)
    Cc = ch // 4
    assert N == (h // 2) * (w // 2)
    out = torch.empty((B, Cc, h, w), device=x.device, dtype=x.dtype)
    L.check(_lib.lx_pack_latents(x.data_ptr(), out.data_ptr(), B, Cc, h, w, x.element_size(), 1, _stream()),
            "lx_pack_latents")
    return out
