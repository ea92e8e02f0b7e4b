"""Tensor-level wrappers over the C ABI (include/loongx_b200.h).

Each function takes torch CUDA tensors, extracts raw pointers / strides and calls the native entry point on the
current torch CUDA stream.  No arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib as L


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda, "loongx_b200 ops need CUDA tensors (there is no CPU fallback)"
    return t.data_ptr()


def _stream() -> int:
    return L.current_stream()


def make_tile_meta(batch: int, n_txt: int, n_img: int, n_cond: int, device) -> torch.Tensor:
    """Per-128-row tile metadata for the stream-major row layout [txt(B*Nt) | img(B*Ni) | cond(B*Nc)].

    Integer layout work, bit-exact by construction: tile -> (stream, batch, batch*S + offset in [txt|img|cond]).
    """
    S = n_txt + n_img + n_cond
    rows = []
    for stream, (n, off) in enumerate(((n_txt, 0), (n_img, n_txt), (n_cond, n_txt + n_img))):
        assert n % 128 == 0, f"stream length {n} must be a multiple of 128"
        for b in range(batch):
            for t in range(n // 128):
                rows.append((stream, b, b * S + off + t * 128, 0))
    return torch.tensor(rows, dtype=torch.int32, device=device).reshape(-1, 4).contiguous()


def gemm(
    A: torch.Tensor,
    W: torch.Tensor,
    bias: Optional[torch.Tensor],
    out: Optional[torch.Tensor] = None,
    mode: int = L.EPI_BIAS,
    *,
    K: Optional[int] = None,
    col_offset: int = 0,
    n_split: Optional[int] = None,
    seg1: Optional[tuple] = None,  # (mode, out, col_offset)
    out2: Optional[tuple] = None,  # (tensor, col_offset2): second output of the EPI_BIAS_GELU_DUAL segment
    groups: Optional[Sequence[tuple]] = None,  # extra row groups: (W, bias, m_begin)
    tile_meta: Optional[torch.Tensor] = None,
    residual: Optional[torch.Tensor] = None,
    gate: Optional[Sequence[Optional[torch.Tensor]]] = None,  # per stream [B, *] views (row = batch)
    qkv: Optional[tuple] = None,  # (q, k, v) each [B, H, S, 128]
    qkv_pre: Optional[torch.Tensor] = None,  # EPI_QKV: also keep the pre-norm projection [M, >= 3 H 128] (training)
    rms_q: Optional[Sequence[Optional[torch.Tensor]]] = None,
    rms_k: Optional[Sequence[Optional[torch.Tensor]]] = None,
    rope: Optional[torch.Tensor] = None,
    rms_eps: float = 1e-6,
    tile_n: int = 0,
    w_dynamic: bool = False,  # W is an activation written by an earlier kernel (no W prefetch ahead of the PDL wait)
) -> None:
    """C = epilogue(A[M,K] @ W[N,K]^T); A / W are 2-D bf16 views with unit inner stride.  `groups` adds row groups
    (rows >= m_begin use that weight panel / bias instead)."""
    assert A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16
    assert A.stride(1) == 1 and W.stride(1) == 1
    d = L.GemmDesc()
    d.A, d.lda = _ptr(A), A.stride(0)
    d.M, d.N = A.shape[0], W.shape[0]
    kk = K if K is not None else A.shape[1]
    assert W.shape[1] >= kk and A.shape[1] >= kk
    panels = [(W, bias, 0)] + list(groups or [])
    d.n_groups = len(panels)
    for i, (w_i, b_i, m_i) in enumerate(panels):
        assert w_i.dtype == torch.bfloat16 and w_i.stride(1) == 1 and w_i.shape[0] == d.N
        if b_i is not None:
            assert b_i.dtype == torch.float32
        d.group[i].W, d.group[i].ldw, d.group[i].bias = _ptr(w_i), w_i.stride(0), _ptr(b_i)
        d.group[i].K, d.group[i].m_begin = kk, m_i
    d.n_split = n_split if n_split is not None else d.N
    d.seg[0].mode = mode
    d.seg[0].col_offset = col_offset
    if out is not None:
        assert out.stride(-1) == 1
        d.seg[0].out, d.seg[0].ldo = _ptr(out), out.stride(0)
    if seg1 is not None:
        m1, o1, c1 = seg1
        d.seg[1].mode, d.seg[1].out, d.seg[1].ldo, d.seg[1].col_offset = m1, _ptr(o1), o1.stride(0), c1
    if out2 is not None:
        o2, c2 = out2
        assert o2.dtype == torch.bfloat16 and o2.stride(-1) == 1
        d.out2, d.ldo2, d.col_offset2 = _ptr(o2), o2.stride(0), c2
    d.tile_meta = _ptr(tile_meta)
    if residual is not None:
        d.residual, d.ldr = _ptr(residual), residual.stride(0)
    if gate is not None:
        for i, g in enumerate(gate):
            if g is not None:
                assert g.dtype == torch.bfloat16 and g.stride(-1) == 1
                d.gate[i] = _ptr(g)
                d.gate_stride[i] = g.stride(0) if g.dim() > 1 else 0
    if qkv is not None:
        q, k, v = qkv
        assert q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
        d.q, d.k, d.v = _ptr(q), _ptr(k), _ptr(v)
        d.heads, d.seq_total = q.shape[1], q.shape[2]
        for i in range(3):
            if rms_q is not None and rms_q[i] is not None:
                d.rms_q[i] = _ptr(rms_q[i])
            if rms_k is not None and rms_k[i] is not None:
                d.rms_k[i] = _ptr(rms_k[i])
        d.rope = _ptr(rope)
        if qkv_pre is not None:
            assert qkv_pre.dtype == torch.bfloat16 and qkv_pre.stride(-1) == 1
            d.qkv_pre, d.ld_qkv_pre = _ptr(qkv_pre), qkv_pre.stride(0)
    d.rms_eps = rms_eps
    d.tile_n = tile_n
    d.w_dynamic = int(bool(w_dynamic))
    L.check(L.lib.lx_gemm_bf16(C.byref(d), _stream()), "lx_gemm_bf16")


def make_out_row_base(batch: int, n_txt: int, n_img: int, n_cond: int, device) -> torch.Tensor:
    """Inverse of make_tile_meta: (batch, sequence tile) -> first row in the stream-major activation layout."""
    base = []
    for b in range(batch):
        for n, off in ((n_txt, 0), (n_img, batch * n_txt), (n_cond, batch * (n_txt + n_img))):
            for t in range(n // 128):
                base.append(off + b * n + t * 128)
    return torch.tensor(base, dtype=torch.int32, device=device)


def _set_pads(d, S: int, n_cond: int, pads: Optional[Sequence[int]], n_txt: Optional[int]) -> None:
    """pads = (pad_txt, pad_img, pad_cond); n_txt = padded text length (needed to locate the end of the text stream)."""
    if pads is None or not any(pads):
        return
    assert n_txt is not None, "n_txt (padded) is needed to place the stream boundaries"
    d.stream_end[0], d.stream_end[1], d.stream_end[2] = n_txt, S - n_cond, S
    for i in range(3):
        d.pad[i] = int(pads[i])


def attention(q, k, v, out, out_row_base, *, n_cond: int = 0, mask_mode: int = 0, cross_bias: float = 0.0,
              col_offset: int = 0, scale: Optional[float] = None, lse: Optional[torch.Tensor] = None,
              pads: Optional[Sequence[int]] = None, n_txt: Optional[int] = None) -> None:
    """out rows <- softmax(q k^T * scale) v for q,k,v [B,H,S,128] bf16 (see lx_attention)."""
    assert q.dtype == torch.bfloat16 and q.is_contiguous() and k.is_contiguous() and v.is_contiguous()
    B, H, S, Dh = q.shape
    assert Dh == 128
    d = L.AttnDesc()
    d.q, d.k, d.v = _ptr(q), _ptr(k), _ptr(v)
    d.out, d.ldo = _ptr(out), out.stride(0)
    d.out_row_base = _ptr(out_row_base)
    d.col_offset = col_offset
    d.B, d.H, d.S = B, H, S
    d.n_cond, d.mask_mode = n_cond, mask_mode
    d.cross_bias = cross_bias
    d.scale = scale if scale is not None else 1.0 / (Dh ** 0.5)
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.numel() == B * H * S
        d.lse = _ptr(lse)
    _set_pads(d, S, n_cond, pads, n_txt)
    L.check(L.lib.lx_attention(C.byref(d), _stream()), "lx_attention")


def attention_bwd(q, k, v, d_out_heads, lse, delta, dq_f32, dk, dv, *, n_cond: int = 0, mask_mode: int = 0,
                  cross_bias: float = 0.0, scale: Optional[float] = None, pads: Optional[Sequence[int]] = None,
                  n_txt: Optional[int] = None) -> None:
    """dq_f32 (fp32, zeroed by the caller) += dQ; dk, dv (bf16) = dK, dV of lx_attention (see lx_attention_bwd)."""
    B, H, S, Dh = q.shape
    assert Dh == 128 and dq_f32.dtype == torch.float32 and dk.dtype == torch.bfloat16 and dv.dtype == torch.bfloat16
    for t in (q, k, v, d_out_heads, lse, delta, dq_f32, dk, dv):
        assert t.is_cuda and t.is_contiguous()
    d = L.AttnBwdDesc()
    d.q, d.k, d.v, d.d_out = _ptr(q), _ptr(k), _ptr(v), _ptr(d_out_heads)
    d.lse, d.delta, d.dq, d.dk, d.dv = _ptr(lse), _ptr(delta), _ptr(dq_f32), _ptr(dk), _ptr(dv)
    d.B, d.H, d.S, d.n_cond, d.mask_mode = B, H, S, n_cond, mask_mode
    d.cross_bias = cross_bias
    d.scale = scale if scale is not None else 1.0 / (Dh ** 0.5)
    _set_pads(d, S, n_cond, pads, n_txt)
    L.check(L.lib.lx_attention_bwd(C.byref(d), _stream()), "lx_attention_bwd")


def attention_bwd_prep(d_out_rows, out_rows, heads: int, tile_meta, d_out_heads, delta) -> None:
    L.check(L.lib.lx_attention_bwd_prep(_ptr(d_out_rows), d_out_rows.stride(0), _ptr(out_rows), out_rows.stride(0),
                                        d_out_rows.shape[0], heads, _ptr(tile_meta), _ptr(d_out_heads), _ptr(delta),
                                        d_out_heads.shape[2], _stream()), "lx_attention_bwd_prep")
