"""Drop-in for the reference's src/flux/block.py: `attn_forward`, `block_forward`, `single_block_forward`
(block.py:7-176, 179-278, 281-339) at the reference's granularity, on the native kernels.

`self` / `attn` are the block / attention handles of `NativeFluxTransformer` (`transformer.transformer_blocks[i]`,
`.single_transformer_blocks[i]`, `block.attn`) — the objects the reference gets from diffusers.  Tensors come in and go out
in the reference's [B, N, D] layout; inside they are copied into the stream-major row layout of a cached `DitPlan`
(padding ragged streams to 128-token tiles), the block's AdaLN vectors are produced from `temb` / `cond_temb` with the
block's slice of the stacked modulation panel, and ONE native block call runs (`lx_dit_double_block` /
`lx_dit_single_block`, or GEMM + attention + GEMM for `attn_forward`).  Python here is layout plumbing only.
"""
from typing import Any, Dict, Optional

import torch

from loongx_b200 import _lib as L
from loongx_b200 import ops
from loongx_b200.train import ln_modulate  # noqa: F401  (re-exported for callers that want the AdaLN apply alone)

_lib = L.lib
_lib.lx_add_silu_bcast.argtypes = [L.c_void_p, L.c_void_p, L.c_void_p, L.c_int32, L.c_void_p, L.c_int64, L.c_int32, L.c_int32,
                                   L.c_void_p]


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _bf(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).contiguous()


def _plan(block, B, n_txt, n_img, n_cond, model_config):
    tr = block.transformer
    return tr.plan(B, n_txt, n_img, n_cond, 1, model_config or {}, tr.c_factor())


def _load_streams(plan, dst, txt, img, cond):
    """[B, n, C] tensors -> stream-major (padded) rows of `dst` [R, C]."""
    B = plan.B
    if plan.padded:
        dst.zero_()
    rt, ri = B * plan.ntp, B * plan.nip
    dst[:rt].view(B, plan.ntp, -1)[:, :plan.nt].copy_(txt)
    dst[rt:rt + ri].view(B, plan.nip, -1)[:, :plan.ni].copy_(img)
    if cond is not None:
        dst[rt + ri:].view(B, plan.ncp, -1)[:, :plan.nc].copy_(cond)


def _split_rows(plan, rows, dtype):
    B = plan.B
    rt, ri = B * plan.ntp, B * plan.nip
    txt = rows[:rt].view(B, plan.ntp, -1)[:, :plan.nt].to(dtype)
    img = rows[rt:rt + ri].view(B, plan.nip, -1)[:, :plan.ni].to(dtype)
    cond = rows[rt + ri:].view(B, plan.ncp, -1)[:, :plan.nc].to(dtype) if plan.nc else None
    return txt.clone(), img.clone(), (cond.clone() if cond is not None else None)


def _set_rope(plan, image_rotary_emb, cond_rotary_emb):
    """(cos, sin) [n, 128] with every value repeated for the two elements of a rotary pair (FluxPosEmbed) -> the native
    [S, 64, 2] table at the (padded) stream positions."""
    table = plan.buf["rope"]
    table.zero_()
    table[..., 0] = 1.0  # identity rotation on padding tokens / when no embedding is given
    if image_rotary_emb is not None:
        cos, sin = (t.to(table.device, torch.float32) for t in image_rotary_emb)
        assert cos.shape[0] == plan.nt + plan.ni, (cos.shape, plan.nt, plan.ni)
        for src0, n, dst0 in ((0, plan.nt, 0), (plan.nt, plan.ni, plan.ntp)):
            table[dst0:dst0 + n, :, 0] = cos[src0:src0 + n, 0::2]
            table[dst0:dst0 + n, :, 1] = sin[src0:src0 + n, 0::2]
    if cond_rotary_emb is not None and plan.nc:
        cos, sin = (t.to(table.device, torch.float32) for t in cond_rotary_emb)
        d0 = plan.ntp + plan.nip
        table[d0:d0 + plan.nc, :, 0] = cos[:, 0::2]
        table[d0:d0 + plan.nc, :, 1] = sin[:, 0::2]
    plan.has_rope = True


def _modulation(plan, panel, blk, width, emb, out_table, lora: bool):
    """out_table[:, blk*width : (blk+1)*width] = Linear(silu(emb)) with block `blk`'s rows of the stacked AdaLN panel
    (AdaLayerNormZero / -Single); `lora` picks the merged panel (condition stream / latent_lora)."""
    B, D = emb.shape
    silu = plan.buf["silu_c"]
    e = _bf(emb)
    L.check(_lib.lx_add_silu_bcast(e.data_ptr(), None, None, 1, silu.data_ptr(), D, B, D, _stream()), "lx_add_silu_bcast")
    W = panel.w_lora if (lora and panel.w_lora is not None) else panel.w
    rows = slice(blk * width, (blk + 1) * width)
    ops.gemm(silu, W[rows], panel.bias[rows], out_table[:B, rows], L.EPI_BIAS)


def _single_split(n_total: int):
    """single blocks get text and image tokens as one tensor; any split works (same weights and modulation)."""
    nt = max(1, min(128, n_total - 1))
    return nt, n_total - nt


def block_forward(self, hidden_states, encoder_hidden_states, condition_latents, temb, cond_temb, cond_rotary_emb=None,
                  image_rotary_emb=None, model_config: Optional[Dict[str, Any]] = {}):
    """block.py:179-278 -> (encoder_hidden_states, hidden_states, condition_latents or None)."""
    model_config = model_config or {}
    if self.single:
        raise TypeError("block_forward needs a double-stream block (transformer.transformer_blocks[i])")
    use_cond = condition_latents is not None
    B, ni, D = hidden_states.shape
    nt = encoder_hidden_states.shape[1]
    nc = condition_latents.shape[1] if use_cond else 0
    plan = _plan(self, B, nt, ni, nc, model_config)
    named = self.transformer.weights.named
    ll = bool(model_config.get("latent_lora", False))
    i = self.index
    _load_streams(plan, plan.buf["X"], _bf(encoder_hidden_states), _bf(hidden_states), _bf(condition_latents) if use_cond else None)
    _set_rope(plan, image_rotary_emb, cond_rotary_emb if use_cond else None)
    _modulation(plan, named["mod_img"], i, 6 * D, temb, plan.buf["mod_img"], ll)
    _modulation(plan, named["mod_txt"], i, 6 * D, temb, plan.buf["mod_txt"], False)
    if use_cond:
        _modulation(plan, named["mod_img"], i, 6 * D, cond_temb, plan.buf["mod_cond_img"], True)
    plan.double_block(0, i)
    txt, img, cond = _split_rows(plan, plan.buf["X"], hidden_states.dtype)
    return txt, img, cond if use_cond else None


def single_block_forward(self, hidden_states, temb, image_rotary_emb=None, condition_latents=None, cond_temb=None,
                         cond_rotary_emb=None, model_config: Optional[Dict[str, Any]] = {}):
    """block.py:281-339 -> hidden_states, or (hidden_states, condition_latents) with a condition."""
    model_config = model_config or {}
    if not self.single:
        raise TypeError("single_block_forward needs a single-stream block (transformer.single_transformer_blocks[i])")
    use_cond = condition_latents is not None
    B, n_total, D = hidden_states.shape
    nt, ni = _single_split(n_total)
    nc = condition_latents.shape[1] if use_cond else 0
    plan = _plan(self, B, nt, ni, nc, model_config)
    named = self.transformer.weights.named
    ll = bool(model_config.get("latent_lora", False))
    i = self.index
    hs = _bf(hidden_states)
    _load_streams(plan, plan.buf["X"], hs[:, :nt], hs[:, nt:], _bf(condition_latents) if use_cond else None)
    _set_rope(plan, image_rotary_emb, cond_rotary_emb if use_cond else None)
    _modulation(plan, named["mod_single"], i, 3 * D, temb, plan.buf["mod_single"], ll)
    if use_cond:
        _modulation(plan, named["mod_single"], i, 3 * D, cond_temb, plan.buf["mod_cond_single"], True)
    plan.single_block(0, i)
    txt, img, cond = _split_rows(plan, plan.buf["X"], hidden_states.dtype)
    out = torch.cat([txt, img], dim=1)
    return (out, cond) if use_cond else out


def attn_forward(attn, hidden_states, encoder_hidden_states=None, condition_latents=None, attention_mask=None,
                 image_rotary_emb=None, cond_rotary_emb=None, model_config: Optional[Dict[str, Any]] = {}):
    """block.py:7-176: q/k/v projections (+ context projections), per-head RMSNorm, RoPE, joint attention with the block
    masks / c_factor bias, output projections (double-stream blocks only).  Inputs are the already modulated streams."""
    model_config = model_config or {}
    if attention_mask is not None:
        raise NotImplementedError("attn_forward builds its own mask from model_config like the reference (block.py:106-128)")
    block = attn.block
    named = block.transformer.weights.named
    use_cond = condition_latents is not None
    double = encoder_hidden_states is not None
    if double == block.single:
        raise TypeError("encoder_hidden_states goes with double-stream blocks only (block.py:43-44)")
    B, n_h, D = hidden_states.shape
    if double:
        nt, ni = encoder_hidden_states.shape[1], n_h
        txt_in, img_in = _bf(encoder_hidden_states), _bf(hidden_states)
    else:
        nt, ni = _single_split(n_h)
        hs = _bf(hidden_states)
        txt_in, img_in = hs[:, :nt], hs[:, nt:]
    nc = condition_latents.shape[1] if use_cond else 0
    plan = _plan(block, B, nt, ni, nc, model_config)
    b, p = plan.buf, plan.plan
    ll = bool(model_config.get("latent_lora", False))
    i = block.index
    H = block.transformer.cfg.num_attention_heads
    _load_streams(plan, b["XN"], txt_in, img_in, _bf(condition_latents) if use_cond else None)
    _set_rope(plan, image_rotary_emb, cond_rotary_emb if use_cond else None)
    rt, ri = B * plan.ntp, B * plan.nip
    pick = lambda pl, lora: pl.w_lora if (lora and pl.w_lora is not None) else pl.w  # noqa: E731
    if double:
        qkv, ctx = named[f"double.{i}.qkv"], named[f"double.{i}.qkv_ctx"]
        nq, nk = named[f"double.{i}.norm_q"], named[f"double.{i}.norm_k"]
        naq, nak = named[f"double.{i}.norm_added_q"], named[f"double.{i}.norm_added_k"]
        W0, b0 = ctx.w, ctx.bias
        groups = [(pick(qkv, ll), qkv.bias, rt)] + ([(pick(qkv, True), qkv.bias, rt + ri)] if use_cond else [])
        rq, rk = [naq, nq, nq], [nak, nk, nk]
    else:
        pl = named[f"single.{i}.qkv_mlp"]
        nq, nk = named[f"single.{i}.norm_q"], named[f"single.{i}.norm_k"]
        W0, b0 = pick(pl, ll)[:3 * D], pl.bias[:3 * D]
        groups = [(pick(pl, True)[:3 * D], pl.bias[:3 * D], rt + ri)] if use_cond else []
        rq, rk = [nq, nq, nq], [nk, nk, nk]
    ops.gemm(b["XN"], W0, b0, None, L.EPI_QKV, groups=groups, tile_meta=b["tile_meta"], qkv=(b["Q"], b["K"], b["V"]),
             rms_q=rq, rms_k=rk, rope=b["rope"])
    att = b["scratch"].view(-1)[:b["X"].numel()].view_as(b["X"])  # [R, D] rows
    pads = (plan.ntp - plan.nt, plan.nip - plan.ni, plan.ncp - plan.nc)
    ops.attention(b["Q"], b["K"], b["V"], att, b["out_row_base"], n_cond=plan.ncp, mask_mode=p.mask_mode,
                  cross_bias=p.cross_bias, pads=pads, n_txt=plan.ntp)
    dt = hidden_states.dtype
    if not double:  # block.py:168-176
        txt, img, cond = _split_rows(plan, att, dt)
        out = torch.cat([txt, img], dim=1)
        return (out, cond) if use_cond else out
    out_l, out_c = named[f"double.{i}.out"], named[f"double.{i}.out_ctx"]
    groups = [(pick(out_l, ll), out_l.bias, rt)] + ([(pick(out_l, True), out_l.bias, rt + ri)] if use_cond else [])
    ops.gemm(att, out_c.w, out_c.bias, b["X"], L.EPI_BIAS, groups=groups)
    txt, img, cond = _split_rows(plan, b["X"], dt)
    return (img, txt, cond) if use_cond else (img, txt)  # block.py:163-167: (hidden, encoder, condition)
