"""Training-mode forward (keeps what the backward reads) and explicit backward of the CS3 encoders / DGF fusion, on the
native fp32 kernels of csrc/cs3_dgf_bwd.cu.  No torch arithmetic: tensors are allocated with torch, every operation is a
kernel of the C ABI (include/loongx_b200.h, "Backward of CS3 / DGF").

Reference: OminiModel.step (src/train/model.py:656-701) runs the encoders, fuse_eeg / fuse_fnirs, the DUANs and fusion3/4
inside the autograd graph, so `loss.backward()` leaves a gradient on every one of their parameters and Lightning's DDP
all-reduces them with the LoRA factors (train.py:181-183).  Here the gradients accumulate into caller-provided fp32 views
(`grads[param]`, one flat bucket in practice) so that ONE all-reduce covers them together with the LoRA gradients.

Dropout(0.3) of the projection MLPs (model.py:64,68) is active when `training` is true: a counter-based mask keyed by
(seed, layer), identical in forward and backward.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from . import cs3
from .cs3 import DuanWeights, _f32, _p, _st

c_void_p, c_int32, c_int64, c_float, c_uint64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint64
_lib = L.lib


class SgemmExDesc(C.Structure):
    _fields_ = [("A", c_void_p), ("lda", c_int64), ("a_bstride", c_int64), ("Bm", c_void_p), ("ldb", c_int64),
                ("b_bstride", c_int64), ("C", c_void_p), ("ldc", c_int64), ("c_bstride", c_int64), ("M", c_int32),
                ("N", c_int32), ("K", c_int32), ("batch", c_int32), ("trans_a", c_int32), ("trans_b", c_int32),
                ("reduce_batch", c_int32), ("reserved", c_int32), ("alpha", c_float), ("beta", c_float)]


_lib.lx_sgemm_ex.argtypes = [C.POINTER(SgemmExDesc), c_void_p]
_lib.lx_sum_rows_f32.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_void_p]
_lib.lx_sum_last_f32.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int32, c_void_p]
_lib.lx_ln_relu_rows_bwd.argtypes = [c_void_p] * 7 + [c_int32, c_int32, c_float, c_void_p]
_lib.lx_dropout_f32.argtypes = [c_void_p, c_void_p, c_int64, c_float, c_uint64, c_void_p]
_lib.lx_token_linear_bwd.argtypes = [c_void_p] * 6 + [c_int32, c_int32, c_int32, c_int64, c_void_p]
_lib.lx_adaptive_pool_bwd.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int64, c_int32, c_int32,
                                      c_int32, c_void_p]
_lib.lx_channel_ln_bwd.argtypes = [c_void_p] * 6 + [c_int32, c_int32, c_int32, c_float, c_void_p]
_lib.lx_gelu_erf_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
_lib.lx_s4_conv.argtypes = [c_void_p] * 5 + [c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p]
_lib.lx_s4_conv_wgrad.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]
_lib.lx_s4_kernel_gen_bwd.argtypes = [c_void_p] * 11 + [c_int32, c_int32, c_int32, c_void_p]
_lib.lx_duan_backward.argtypes = [C.POINTER(DuanWeights), C.POINTER(DuanWeights), c_void_p, c_void_p, c_void_p, c_int64,
                                  c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                  c_void_p]
_lib.lx_skinny_xw_bf16.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32,
                                   c_void_p]
_lib.lx_silu_bwd_sum.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_void_p]
_lib.lx_silu_bwd_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]

Grads = Dict[nn.Parameter, torch.Tensor]  # parameter -> fp32 gradient view (complex parameters: view_as_real layout)


def _z(*shape, dev):
    return torch.zeros(shape, device=dev, dtype=torch.float32)


def _e(*shape, dev):
    return torch.empty(shape, device=dev, dtype=torch.float32)


# ---------------------------------------------------------------------------------------------------------------
# thin op wrappers
# ---------------------------------------------------------------------------------------------------------------
def sgemm_ex(A, Bm, Cm, M, N, K, *, lda, ldb, ldc, trans_a=0, trans_b=0, batch=1, a_bs=0, b_bs=0, c_bs=0, reduce=0,
             alpha=1.0, beta=0.0):
    d = SgemmExDesc()
    d.A, d.lda, d.a_bstride = A if isinstance(A, int) else A.data_ptr(), lda, a_bs
    d.Bm, d.ldb, d.b_bstride = Bm if isinstance(Bm, int) else Bm.data_ptr(), ldb, b_bs
    d.C, d.ldc, d.c_bstride = Cm if isinstance(Cm, int) else Cm.data_ptr(), ldc, c_bs
    d.M, d.N, d.K, d.batch = M, N, K, batch
    d.trans_a, d.trans_b, d.reduce_batch = trans_a, trans_b, reduce
    d.alpha, d.beta = alpha, beta
    L.check(_lib.lx_sgemm_ex(C.byref(d), _st()), "lx_sgemm_ex")


def sum_rows(x: torch.Tensor, out: torch.Tensor, accumulate=True):
    """out[j] (+)= sum_r x[r, j] for a 2-D (row-strided) x."""
    L.check(_lib.lx_sum_rows_f32(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], out.data_ptr(), int(accumulate), _st()),
            "lx_sum_rows_f32")


def sum_last(a: torch.Tensor, out: torch.Tensor, b: Optional[torch.Tensor] = None, accumulate=True):
    """out[c] (+)= sum_{b,l} a[b,c,l] (* b[b,c,l]) for contiguous [B, C, L]."""
    B, Cc, Ln = a.shape
    L.check(_lib.lx_sum_last_f32(a.data_ptr(), _p(b), B, Cc, Ln, out.data_ptr(), int(accumulate), _st()), "lx_sum_last_f32")


def dropout(x: torch.Tensor, p: float, seed: int) -> torch.Tensor:
    y = torch.empty_like(x)
    L.check(_lib.lx_dropout_f32(x.data_ptr(), y.data_ptr(), x.numel(), p, seed & 0xFFFFFFFFFFFFFFFF, _st()), "lx_dropout_f32")
    return y


def linear_rows_bwd(lin: nn.Linear, x: torch.Tensor, dy: torch.Tensor, grads: Grads, need_dx=True) -> Optional[torch.Tensor]:
    """y = x W^T + b for x [B, n_in], dy [B, n_out]: dW += dy^T x, db += sum_b dy, -> dx = dy W."""
    B, n_in = x.shape
    n_out = dy.shape[1]
    sgemm_ex(dy, x, grads[lin.weight], n_out, n_in, B, lda=dy.stride(0), ldb=x.stride(0), ldc=n_in, trans_a=1, beta=1.0)
    sum_rows(dy, grads[lin.bias])
    if not need_dx:
        return None
    dx = _e(B, n_in, dev=x.device)
    sgemm_ex(dy, lin.weight, dx, B, n_in, n_out, lda=dy.stride(0), ldb=n_in, ldc=n_in)
    return dx


def channel_linear_bwd(lin: nn.Linear, x: torch.Tensor, dy: torch.Tensor, grads: Grads, need_dx=True):
    """y[b] = W x[b] + bias for channel-major x [B, d_in, L], dy [B, d_out, L]."""
    B, d_in, Ln = x.shape
    d_out = dy.shape[1]
    sgemm_ex(dy, x, grads[lin.weight], d_out, d_in, Ln, lda=Ln, ldb=Ln, ldc=d_in, trans_b=1, batch=B, a_bs=d_out * Ln,
             b_bs=d_in * Ln, reduce=1, beta=1.0)
    sum_last(dy, grads[lin.bias])
    if not need_dx:
        return None
    dx = _e(B, d_in, Ln, dev=x.device)
    sgemm_ex(lin.weight, dy, dx, d_in, Ln, d_out, lda=d_in, ldb=Ln, ldc=Ln, trans_a=1, batch=B, b_bs=d_out * Ln,
             c_bs=d_in * Ln)
    return dx


def token_axis_linear_bwd(lin: nn.Linear, x: torch.Tensor, dout: torch.Tensor, grads: Grads, dx_rows: Optional[slice] = None):
    """out[b, j, :] = sum_i W[j, i] x[b, i, :] + bias[j] (cs3.token_axis_linear): dW, db +=; returns dx restricted to
    the input rows `dx_rows` (None = all, False = no input gradient)."""
    B, n_in, Dm = x.shape
    n_out = dout.shape[1]
    sgemm_ex(dout, x, grads[lin.weight], n_out, n_in, Dm, lda=Dm, ldb=Dm, ldc=n_in, trans_b=1, batch=B, a_bs=n_out * Dm,
             b_bs=n_in * Dm, reduce=1, beta=1.0)
    sum_last(dout, grads[lin.bias])
    if dx_rows is False:
        return None
    r0, r1 = (0, n_in) if dx_rows is None else (dx_rows.start, dx_rows.stop)
    m = r1 - r0
    dx = _e(B, m, Dm, dev=x.device)
    # dx[b, i, :] = sum_j W[j, r0 + i] dout[b, j, :]: op(A) = W[:, r0:r1]^T, stored [n_out (K), n_in] at column offset r0
    sgemm_ex(lin.weight.data_ptr() + 4 * r0, dout, dx, m, Dm, n_out, lda=n_in, ldb=Dm, ldc=Dm, trans_a=1, batch=B,
             b_bs=n_out * Dm, c_bs=m * Dm)
    return dx


# ---------------------------------------------------------------------------------------------------------------
# S4 model (channel-major activations [B, d, L])
# ---------------------------------------------------------------------------------------------------------------
def s4model_forward_train(m: cs3.S4Model, x: torch.Tensor):
    x = _f32(x)
    B, d_in, Ln = x.shape
    dev = x.device

    def chan_linear(inp, lin, residual=None, norm=None):
        out = _e(B, lin.out_features, Ln, dev=dev)
        L.check(_lib.lx_channel_linear(inp.data_ptr(), lin.weight.data_ptr(), lin.bias.data_ptr(), _p(residual),
                                       _p(norm.weight) if norm is not None else None,
                                       _p(norm.bias) if norm is not None else None, out.data_ptr(), B, lin.in_features,
                                       lin.out_features, Ln, norm.eps if norm is not None else 0.0, _st()), "lx_channel_linear")
        return out

    h = chan_linear(x, m.encoder)
    saved = []
    for blk in m.blocks:
        K = blk.s4.kernel(Ln)
        g, pre = torch.empty_like(h), torch.empty_like(h)
        L.check(_lib.lx_s4_conv(h.data_ptr(), K.data_ptr(), blk.s4.D.data_ptr(), g.data_ptr(), pre.data_ptr(), B, h.shape[1],
                                Ln, 0, 1, _st()), "lx_s4_conv")
        hn = chan_linear(g, blk.linear, residual=h, norm=blk.norm)
        saved.append((h, pre, g, K))
        h = hn
    out = chan_linear(h, m.decoder)
    return out, dict(x=x, blocks=saved, h_last=h, chan_linear=chan_linear)


def s4model_backward(m: cs3.S4Model, ctx, dout: torch.Tensor, grads: Grads) -> None:
    """dout [B, d_out, L] (channel-major) -> accumulates every parameter gradient of the S4Model (the raw signal needs none)."""
    x, dev = ctx["x"], dout.device
    B, _, Ln = x.shape
    dh = channel_linear_bwd(m.decoder, ctx["h_last"], dout.contiguous(), grads)
    vr = torch.view_as_real
    for blk, (h_in, pre, g, K) in zip(reversed(list(m.blocks)), reversed(ctx["blocks"])):
        d = h_in.shape[1]
        s4 = blk.s4
        v = ctx["chan_linear"](g, blk.linear, residual=h_in)  # pre-LayerNorm sum, recomputed
        dv = torch.empty_like(v)
        L.check(_lib.lx_channel_ln_bwd(v.data_ptr(), blk.norm.weight.data_ptr(), dh.data_ptr(), dv.data_ptr(),
                                       grads[blk.norm.weight].data_ptr(), grads[blk.norm.bias].data_ptr(), B, d, Ln,
                                       blk.norm.eps, _st()), "lx_channel_ln_bwd")
        dg = channel_linear_bwd(blk.linear, g, dv, grads)
        ds = torch.empty_like(dg)
        L.check(_lib.lx_gelu_erf_bwd(pre.data_ptr(), dg.data_ptr(), ds.data_ptr(), ds.numel(), _st()), "lx_gelu_erf_bwd")
        sum_last(ds, grads[s4.D].view(-1), b=h_in)
        dK = _e(d, Ln, dev=dev)
        L.check(_lib.lx_s4_conv_wgrad(ds.data_ptr(), h_in.data_ptr(), dK.data_ptr(), B, d, Ln, 0, _st()), "lx_s4_conv_wgrad")
        dconv = torch.empty_like(ds)
        L.check(_lib.lx_s4_conv(ds.data_ptr(), K.data_ptr(), s4.D.data_ptr(), dconv.data_ptr(), None, B, d, Ln, 1, 0, _st()),
                "lx_s4_conv")
        ws = torch.empty((72 * d * Ln,), device=dev, dtype=torch.uint8)
        L.check(_lib.lx_s4_kernel_gen_bwd(vr(s4.lambda_).data_ptr(), vr(s4.p).data_ptr(), vr(s4.q).data_ptr(),
                                          vr(s4.B.detach()).data_ptr(), vr(s4.Ct.detach()).data_ptr(), s4.log_step.data_ptr(),
                                          dK.data_ptr(), grads[s4.B].data_ptr(), grads[s4.Ct].data_ptr(),
                                          grads[s4.log_step].data_ptr(), ws.data_ptr(), d, s4.n, Ln, _st()),
                "lx_s4_kernel_gen_bwd")
        _axpy(dv, dconv)  # residual path + convolution path
        dh = dv
    channel_linear_bwd(m.encoder, x, dh, grads, need_dx=False)


def _axpy(y: torch.Tensor, x: torch.Tensor) -> None:
    """y += x (the one-row case of the accumulate-rows kernel)."""
    n = y.numel()
    L.check(_lib.lx_sum_rows_f32(x.data_ptr(), n, 1, n, y.data_ptr(), 1, _st()), "lx_sum_rows_f32")


# ---------------------------------------------------------------------------------------------------------------
# projection MLP (model.py:60-72) and encoders
# ---------------------------------------------------------------------------------------------------------------
DROP_P = 0.3


def projection_forward_train(proj: nn.Sequential, feat: torch.Tensor, training: bool, seed: int):
    x1 = cs3.gemv(proj[1].weight, proj[1].bias, feat)
    h1 = cs3.ln_relu(x1, proj[2].weight, proj[2].bias, proj[2].eps)
    h1 = dropout(h1, DROP_P, seed * 4 + 1) if training else h1
    x2 = cs3.gemv(proj[5].weight, proj[5].bias, h1)
    h2 = cs3.ln_relu(x2, proj[6].weight, proj[6].bias, proj[6].eps)
    h2 = dropout(h2, DROP_P, seed * 4 + 2) if training else h2
    ctx = dict(feat=feat, x1=x1, h1=h1, x2=x2, h2=h2, training=training, seed=seed)
    if len(proj) <= 9:
        return h2, ctx
    B = h2.shape[0]
    lin = proj[10]
    out = _e(B, 512, lin.out_features, dev=h2.device)
    L.check(_lib.lx_token_linear(h2.data_ptr(), lin.weight.data_ptr(), lin.bias.data_ptr(), out.data_ptr(), B, 512,
                                 lin.out_features, out.stride(0), _st()), "lx_token_linear")
    return out, ctx


def projection_backward(proj: nn.Sequential, ctx, dout: torch.Tensor, grads: Grads) -> torch.Tensor:
    """-> d feat [B, d_in]."""
    dev = dout.device
    B = ctx["feat"].shape[0]
    if len(proj) > 9:
        lin = proj[10]
        dout = dout.contiguous()
        dh2 = _e(B, 512 * 8, dev=dev)
        L.check(_lib.lx_token_linear_bwd(ctx["h2"].data_ptr(), lin.weight.data_ptr(), dout.data_ptr(), dh2.data_ptr(),
                                         grads[lin.weight].data_ptr(), grads[lin.bias].data_ptr(), B, 512, lin.out_features,
                                         dout.stride(0), _st()), "lx_token_linear_bwd")
    else:
        dh2 = dout.contiguous()

    def ln_relu_bwd(x, norm, dy):
        dx = torch.empty_like(x)
        L.check(_lib.lx_ln_relu_rows_bwd(x.data_ptr(), norm.weight.data_ptr(), norm.bias.data_ptr(), dy.data_ptr(), dx.data_ptr(),
                                         grads[norm.weight].data_ptr(), grads[norm.bias].data_ptr(), x.shape[0], x.shape[1],
                                         norm.eps, _st()), "lx_ln_relu_rows_bwd")
        return dx

    if ctx["training"]:
        dh2 = dropout(dh2, DROP_P, ctx["seed"] * 4 + 2)
    dx2 = ln_relu_bwd(ctx["x2"], proj[6], dh2)
    dh1 = linear_rows_bwd(proj[5], ctx["h1"], dx2, grads)
    if ctx["training"]:
        dh1 = dropout(dh1, DROP_P, ctx["seed"] * 4 + 1)
    dx1 = ln_relu_bwd(ctx["x1"], proj[2], dh1)
    return linear_rows_bwd(proj[1], ctx["feat"], dx1, grads)


def _pool_bwd(dfeat, dz, O, cs, is_, off):
    B, Cc, Ln = dz.shape
    L.check(_lib.lx_adaptive_pool_bwd(dfeat.data_ptr(), dz.data_ptr(), B, Cc, Ln, O, dfeat.stride(0), cs, is_, off, _st()),
            "lx_adaptive_pool_bwd")


def encoder_forward_train(enc: nn.Module, x: torch.Tensor, training: bool, seed: int):
    """The forward of cs3.EEGEncoder / _SmallEncoder with the intermediates kept (model.py:74-134, 186-204)."""
    x = _f32(x)
    B = x.shape[0]
    dev = x.device
    if isinstance(enc, cs3.EEGEncoder):
        feat = _e(B, 4 * 4096, dev=dev)
        z1, c1 = s4model_forward_train(enc.s41, x)
        cs3.adaptive_pool(z1, feat, 4, 1, 4096, 0)
        z2, c2 = s4model_forward_train(enc.s42, x)
        cs3.adaptive_pool(z2, feat, 64, 4096, 1, 64 + 3968)
        cs3.adaptive_pool_multi(x, feat, enc.fpp.output_sizes, 4096, 64)
        out, pc = projection_forward_train(enc.projection, feat, training, seed)
        return out, dict(kind="eeg", s41=c1, s42=c2, proj=pc, z1=z1.shape, z2=z2.shape)
    tot = sum(enc.fpp.output_sizes)
    feat = _e(B, enc.ch * enc.pool_size + enc.ch * tot, dev=dev)
    z, c = s4model_forward_train(enc.s4, x)
    cs3.adaptive_pool(z, feat, enc.pool_size, enc.pool_size, 1, 0)
    cs3.adaptive_pool_multi(x, feat, enc.fpp.output_sizes, tot, enc.ch * enc.pool_size)
    out, pc = projection_forward_train(enc.projection, feat, training, seed)
    return out, dict(kind="small", s4=c, proj=pc, z=z.shape)


def encoder_backward(enc: nn.Module, ctx, dout: torch.Tensor, grads: Grads) -> None:
    dev = dout.device
    dfeat = projection_backward(enc.projection, ctx["proj"], dout, grads)
    if ctx["kind"] == "eeg":
        dz1 = _z(*ctx["z1"], dev=dev)
        _pool_bwd(dfeat, dz1, 4, 1, 4096, 0)
        s4model_backward(enc.s41, ctx["s41"], dz1, grads)
        dz2 = _z(*ctx["z2"], dev=dev)
        _pool_bwd(dfeat, dz2, 64, 4096, 1, 64 + 3968)
        s4model_backward(enc.s42, ctx["s42"], dz2, grads)
        return
    dz = _z(*ctx["z"], dev=dev)
    _pool_bwd(dfeat, dz, enc.pool_size, enc.pool_size, 1, 0)
    s4model_backward(enc.s4, ctx["s4"], dz, grads)


# ---------------------------------------------------------------------------------------------------------------
# DUAN
# ---------------------------------------------------------------------------------------------------------------
def _duan_w(mod: cs3.DUAN, src) -> DuanWeights:
    """DuanWeights over the module's parameters (src = lambda p: p) or over their gradient views (src = grads.get)."""
    w = DuanWeights()
    w.gate_w1, w.gate_b1 = src(mod.gate[0].weight).data_ptr(), src(mod.gate[0].bias).data_ptr()
    w.gate_w2, w.gate_b2 = src(mod.gate[2].weight).data_ptr(), src(mod.gate[2].bias).data_ptr()
    w.mlp_w1, w.mlp_b1 = src(mod.mlp[0].weight).data_ptr(), src(mod.mlp[0].bias).data_ptr()
    w.mlp_w2, w.mlp_b2 = src(mod.mlp[2].weight).data_ptr(), src(mod.mlp[2].bias).data_ptr()
    w.hidden, w.eps = mod.hidden_dim, mod.eps
    return w


def duan_forward_train(mod: cs3.DUAN, x: torch.Tensor, c: torch.Tensor, out: Optional[torch.Tensor] = None):
    """DUAN.forward keeping the kernel workspace (statistics, gate mean, FiLM, mask, hidden activations) for the backward."""
    x, c = _f32(x), _f32(c)
    B, Cc, Ln = x.shape
    if out is None:
        out = _e(B, Cc, Ln, dev=x.device)
    ws = _e(cs3.duan_workspace_floats(B, Cc, mod.hidden_dim, Ln), dev=x.device)
    w = _duan_w(mod, lambda p: p)
    L.check(_lib.lx_duan_forward(C.byref(w), x.data_ptr(), c.data_ptr(), out.data_ptr(), out.stride(0), B, Cc, Ln,
                                 float(mod.keep_ratio), ws.data_ptr(), _st()), "lx_duan_forward")
    return out, dict(x=x, c=c, ws=ws)


def duan_backward(mod: cs3.DUAN, ctx, dy: torch.Tensor, grads: Grads, dx: Optional[torch.Tensor] = None,
                  dc: Optional[torch.Tensor] = None, acc_dx=False, acc_dc=False) -> None:
    """dy: [B, C, L] view with unit L stride and C-stride L (batch stride free).  dx / dc: contiguous [B, C, L] or None."""
    x, c = ctx["x"], ctx["c"]
    B, Cc, Ln = x.shape
    assert dy.stride(2) == 1 and dy.stride(1) == Ln
    scratch = _e(B * Cc * Ln + B * mod.hidden_dim * Ln + B * (8 * Cc + 2 * mod.hidden_dim), dev=x.device)
    w, dw = _duan_w(mod, lambda p: p), _duan_w(mod, lambda p: grads[p])
    L.check(_lib.lx_duan_backward(C.byref(w), C.byref(dw), x.data_ptr(), c.data_ptr(), dy.data_ptr(), dy.stride(0), _p(dx),
                                  _p(dc), int(acc_dx), int(acc_dc), B, Cc, Ln, ctx["ws"].data_ptr(), scratch.data_ptr(), _st()),
                  "lx_duan_backward")


# ---------------------------------------------------------------------------------------------------------------
# the conditioning of OminiModel.step (model.py:656-701), forward with context + backward
# ---------------------------------------------------------------------------------------------------------------
def trainable_parameters(model: nn.Module) -> List[nn.Parameter]:
    """Every CS3 / DGF parameter of an OminiModel, in nn.Module order (the encoder half of the gradient bucket)."""
    return [p for p in model.parameters()]


def grad_layout(params: List[nn.Parameter]) -> Tuple[List[Tuple[int, int]], int]:
    """(offset, elements) of every parameter's gradient inside the encoder half of the flat bucket, and the total.
    Complex parameters take 2 floats per element (view_as_real layout); every slice starts on a 16-byte boundary (also
    what torch.view_as_complex needs: an even float offset)."""
    out, o = [], 0
    for p in params:
        n = p.numel() * (2 if p.is_complex() else 1)
        out.append((o, n))
        o += (n + 3) // 4 * 4
    return out, o


def grad_views(params: List[nn.Parameter], flat: torch.Tensor) -> Grads:
    """Carve `flat` (fp32) into one gradient view per parameter (what the kernels accumulate into)."""
    layout, total = grad_layout(params)
    assert total <= flat.numel()
    return {p: flat[o:o + n] for p, (o, n) in zip(params, layout)}


def grad_elements(params: List[nn.Parameter]) -> int:
    return grad_layout(params)[1]


def step_conditioning_train(model, prompt_embeds, pooled, eeg, fnirs, ppg, motion, training: bool, seed: int):
    """model.py:656-701 (the *step* fuse order) -> (prompt_embeds, pooled) in the DiT dtype + the backward context."""
    spp = model.spatial_pyramid_pooling
    ctx: dict = dict(fuse=bool(model.fuse_flag))
    pe_b = po_b = None
    if eeg is not None:
        e, ctx["eeg"] = encoder_forward_train(model.eeg_projection, spp(_f32(eeg), model.eeg_fixed_length), training, seed + 11)
        if ppg is not None:
            p, ctx["ppg"] = encoder_forward_train(model.ppg_projection, spp(_f32(ppg), model.ppg_fixed_length), training, seed + 12)
            B, n_tok, Dm = e.shape
            cat = _e(B, 2 * n_tok, Dm, dev=e.device)
            cat[:, :n_tok].copy_(e)
            _, ctx["duan1"] = duan_forward_train(model.duan_norm1, p, e, out=cat[:, n_tok:])
            pe_b = cs3.token_axis_linear(model.fusion1[0], cat)
            ctx["cat1"] = cat
        else:
            pe_b = e
    if fnirs is not None:
        f, ctx["fnirs"] = encoder_forward_train(model.fnirs_projection, spp(_f32(fnirs), model.fnirs_fixed_length), training, seed + 13)
        if motion is not None:
            m, ctx["motion"] = encoder_forward_train(model.motion_projection, spp(_f32(motion), model.motion_fixed_length), training,
                                                     seed + 14)
            B, Dm = f.shape
            cat = _e(B, 2 * Dm, dev=f.device)
            cat[:, :Dm].copy_(f)
            fused, ctx["duan2"] = duan_forward_train(model.duan_norm2, f.unsqueeze(1).contiguous(), m.unsqueeze(1).contiguous())
            cat[:, Dm:].copy_(fused.squeeze(1))
            po_b = cs3.gemv(model.fusion2[0].weight, model.fusion2[0].bias, cat)
            ctx["cat2"] = cat
        else:
            po_b = f
    if pe_b is None or po_b is None:
        raise ValueError("step(): use_brain_condition needs eeg and fnirs (model.py:680-701 reads both embeddings)")
    if not model.fuse_flag:  # model.py:699-701
        return model.to_model_dtype(pe_b), model.to_model_dtype(po_b), ctx
    to32 = lambda t: _f32(t) if t.dtype == torch.float32 else cs3.cast_f32(t.contiguous())  # noqa: E731
    pe32, po32 = to32(prompt_embeds), to32(pooled)
    B, n_tok, Dm = pe32.shape
    cat3 = _e(B, 2 * n_tok, Dm, dev=pe32.device)
    cat3[:, :n_tok].copy_(pe32)
    _, ctx["duan_prompt"] = duan_forward_train(model.duan_norm_prompt, pe_b, pe32, out=cat3[:, n_tok:])
    pe = cs3.token_axis_linear(model.fusion3[0], cat3, residual=pe32)
    ctx["cat3"] = cat3
    Dp = po32.shape[1]
    fp, ctx["duan_pooled"] = duan_forward_train(model.duan_norm_pooled, po_b.unsqueeze(1).contiguous(),
                                                po32.unsqueeze(1).contiguous())
    catp = _e(B, 2 * Dp, 1, dev=po32.device)
    catp[:, :Dp, 0].copy_(po32)
    catp[:, Dp:, 0].copy_(fp.squeeze(1))
    po = cs3.token_axis_linear(model.fusion4[0], catp, residual=po32.unsqueeze(2).contiguous()).squeeze(2)
    ctx["catp"] = catp
    return model.to_model_dtype(pe), model.to_model_dtype(po), ctx


def step_conditioning_backward(model, ctx, d_pe: torch.Tensor, d_po: torch.Tensor, grads: Grads) -> None:
    """d_pe [B, 512, 4096], d_po [B, 768] (fp32): gradients of the loss w.r.t. the conditioned embeddings the DiT received
    -> every CS3 / DGF parameter gradient (+=)."""
    d_pe, d_po = _f32(d_pe), _f32(d_po)
    dev = d_pe.device
    B, n_tok, Dm = d_pe.shape
    Dp = d_po.shape[1]
    if ctx["fuse"]:
        # pe = pe32 + fusion3(cat3): only the DUAN half of cat3 depends on the encoders
        d_fused = token_axis_linear_bwd(model.fusion3[0], ctx["cat3"], d_pe, grads, dx_rows=slice(n_tok, 2 * n_tok))
        d_pe_b = _e(B, n_tok, Dm, dev=dev)
        duan_backward(model.duan_norm_prompt, ctx["duan_prompt"], d_fused, grads, dx=d_pe_b)
        d_fp = token_axis_linear_bwd(model.fusion4[0], ctx["catp"], d_po.unsqueeze(2).contiguous(), grads,
                                     dx_rows=slice(Dp, 2 * Dp))  # [B, Dp, 1]
        d_po_b = _e(B, 1, Dp, dev=dev)
        duan_backward(model.duan_norm_pooled, ctx["duan_pooled"], d_fp.view(B, 1, Dp), grads, dx=d_po_b)
        d_po_b = d_po_b.view(B, Dp)
    else:
        d_pe_b, d_po_b = d_pe, d_po
    # prompt side
    if "ppg" in ctx:
        d_cat = token_axis_linear_bwd(model.fusion1[0], ctx["cat1"], d_pe_b, grads)  # [B, 2 n_tok, Dm]
        d_e = d_cat[:, :n_tok].contiguous()
        d_p = _e(B, n_tok, Dm, dev=dev)
        duan_backward(model.duan_norm1, ctx["duan1"], d_cat[:, n_tok:], grads, dx=d_p, dc=d_e, acc_dc=True)
        encoder_backward(model.ppg_projection, ctx["ppg"], d_p, grads)
        encoder_backward(model.eeg_projection, ctx["eeg"], d_e, grads)
    else:
        encoder_backward(model.eeg_projection, ctx["eeg"], d_pe_b, grads)
    # pooled side
    if "motion" in ctx:
        lin = model.fusion2[0]
        d_cat = linear_rows_bwd(lin, ctx["cat2"], d_po_b.contiguous(), grads)  # [B, 2 Dp]
        d_f = d_cat[:, :Dp].contiguous().view(B, 1, Dp)
        d_m = _e(B, 1, Dp, dev=dev)
        duan_backward(model.duan_norm2, ctx["duan2"], d_cat[:, Dp:].contiguous().view(B, 1, Dp), grads, dx=d_f, dc=d_m,
                      acc_dx=True)
        encoder_backward(model.motion_projection, ctx["motion"], d_m.view(B, Dp), grads)
        encoder_backward(model.fnirs_projection, ctx["fnirs"], d_f.view(B, Dp), grads)
    else:
        encoder_backward(model.fnirs_projection, ctx["fnirs"], d_po_b, grads)
