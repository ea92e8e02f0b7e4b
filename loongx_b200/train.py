"""Training step of the LoongX DiT (reference: OminiModel.step, src/train/model.py:569-729, around tranformer_forward with
`gradient_checkpointing`, transformer.py:138-228): rectified-flow loss and the gradients of the LoRA factors — the
parameters the reference's optimizer trains (`self.trainable_params = self.lora_layers`, model.py:541).

Python here is launch sequencing and pointer plumbing only; every arithmetic step is a native kernel behind the C ABI:

  forward   block.py:179-339 with the epilogue fusions undone where the backward needs an intermediate: tcgen05 GEMMs
            (plain bias epilogue), lx_qkv_post_fwd, the tcgen05 attention kernel, lx_gate_residual_fwd, lx_gelu_fwd,
            lx_ln_modulate.  The residual stream entering each block is checkpointed; the block is recomputed in the
            backward (the reference's torch.utils.checkpoint per block).
  backward  dX of every Linear = the same tcgen05 GEMM against the transposed weight panel (row groups: text rows ->
            *_context^T, image rows -> W^T, condition rows -> (W + sBA)^T); lx_gate_bwd, lx_gelu_bwd, lx_ln_modulate_bwd,
            lx_qkv_post_bwd, lx_lora_grad, lx_flow_mse_loss.
            dQ/dK/dV of the joint attention: lx_attention_bwd (tcgen05, P^T / dS^T resident in tensor memory) from the
            log-sum-exp rows the forward kernel records.  `attn_bwd="library"` swaps in torch's SDPA autograd for A/B
            checks only.

Scope of the gradients: LoRA A / B of every target of train/config/seed_512.yaml:38 — on the condition branch
(`latent_lora=False`, the shipped configuration) or on the image / text+image rows as well (`latent_lora=True`).  The
CS3 / DGF encoders run forward-only (they are not in the reference's optimizer parameter list).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib as L
from . import ops
from .dit import DitPlan, DitWeights, PackedLinear, _stream

c_void_p, c_int32, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float
_P3 = c_void_p * 3
_I3 = c_int64 * 3
_lib = L.lib


class LnModDesc(C.Structure):
    _fields_ = [("x", c_void_p), ("ldx", c_int64), ("out", c_void_p), ("ldo", c_int64), ("rows", c_int32), ("D", c_int32),
                ("tile_meta", c_void_p), ("shift", _P3), ("scale", _P3), ("stride", _I3), ("eps", c_float),
                ("reserved", c_int32)]


_lib.lx_ln_modulate.argtypes = [C.POINTER(LnModDesc), c_void_p]
_lib.lx_gelu_fwd.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p]
_lib.lx_gelu_bwd.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p]
_lib.lx_gate_residual_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, _P3, _I3, c_void_p]
_lib.lx_gate_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, _P3, _I3, _P3, _I3, c_void_p]
_lib.lx_ln_modulate_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, _P3, _I3,
                                    _P3, _P3, _I3, c_float, c_void_p, c_void_p]
_lib.lx_qkv_post_fwd.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, _P3,
                                 _P3, c_void_p, c_float, c_void_p]
_lib.lx_qkv_post_bwd.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                 c_void_p, c_int32, _P3, _P3, c_void_p, c_float, c_void_p]
_lib.lx_qkv_post_bwd_f32dq.argtypes = _lib.lx_qkv_post_bwd.argtypes
_lib.lx_rows_to_heads.argtypes = [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p, c_int32, c_void_p]
_lib.lx_lora_grad.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32,
                              c_int32, c_int32, c_float, c_void_p, c_void_p]
_lib.lx_lora_merge.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_float,
                               c_void_p]
_lib.lx_transpose_bf16.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p]
_lib.lx_flow_noise_mix.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p]
_lib.lx_flow_mse_loss.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_void_p]
_lib.lx_cast.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _vec3(vs) -> Tuple[_P3, _I3]:
    """per-stream [B, n] views (row = batch element) -> (pointers, batch strides)."""
    p, st = _P3(), _I3()
    for i, v in enumerate(vs):
        if v is not None:
            assert v.stride(-1) == 1
            p[i], st[i] = v.data_ptr(), v.stride(0)
    return p, st


# ------------------------------------------------------------------------------------------------------------------
# thin kernel wrappers (pointers + strides only)
# ------------------------------------------------------------------------------------------------------------------
def ln_modulate(x, out, tile_meta, shift, scale, eps=1e-6):
    d = LnModDesc()
    d.x, d.ldx, d.out, d.ldo = x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0)
    d.rows, d.D = x.shape
    d.tile_meta = tile_meta.data_ptr()
    for i in range(3):
        if shift[i] is not None:
            assert shift[i].stride(0) == scale[i].stride(0)
            d.shift[i], d.scale[i], d.stride[i] = shift[i].data_ptr(), scale[i].data_ptr(), shift[i].stride(0)
    d.eps = eps
    L.check(_lib.lx_ln_modulate(C.byref(d), _stream()), "lx_ln_modulate")


def gelu_fwd(pre, out):
    L.check(_lib.lx_gelu_fwd(pre.data_ptr(), pre.stride(0), out.data_ptr(), out.stride(0), pre.shape[0], pre.shape[1],
                             _stream()), "lx_gelu_fwd")


def gelu_bwd(pre, dy, dx):
    L.check(_lib.lx_gelu_bwd(pre.data_ptr(), pre.stride(0), dy.data_ptr(), dy.stride(0), dx.data_ptr(), dx.stride(0),
                             pre.shape[0], pre.shape[1], _stream()), "lx_gelu_bwd")


def gate_residual_fwd(res, y, out, tile_meta, gate):
    assert res.stride(0) == y.stride(0) == out.stride(0)
    gp, gs = _vec3(gate)
    L.check(_lib.lx_gate_residual_fwd(res.data_ptr(), y.data_ptr(), out.data_ptr(), res.stride(0), res.shape[0], res.shape[1],
                                      tile_meta.data_ptr(), gp, gs, _stream()), "lx_gate_residual_fwd")


def add_rows(x, r):
    """x += r, bf16 rows (fp32 add, one rounding)."""
    assert x.shape == r.shape and x.stride(1) == 1 and r.stride(1) == 1
    L.check(_lib.lx_add_rows(x.data_ptr(), x.stride(0), r.data_ptr(), r.stride(0), x.shape[0], x.shape[1], _stream()),
            "lx_add_rows")


def gate_bwd(dout, y, dy, tile_meta, gate, dgate):
    assert dout.stride(0) == y.stride(0) == dy.stride(0)
    gp, gs = _vec3(gate)
    dp, ds = _vec3(dgate)
    L.check(_lib.lx_gate_bwd(dout.data_ptr(), y.data_ptr(), dy.data_ptr(), dout.stride(0), dout.shape[0], dout.shape[1],
                             tile_meta.data_ptr(), gp, gs, dp, ds, _stream()), "lx_gate_bwd")


def ln_modulate_bwd(x, dxn, dres, dx, tile_meta, scale, dscale, dshift, stats, eps=1e-6):
    assert x.stride(0) == dxn.stride(0) == dx.stride(0) and (dres is None or dres.stride(0) == x.stride(0))
    sp, ss = _vec3(scale)
    dsp, dss = _vec3(dscale)
    dhp, dhs = _vec3(dshift)
    for i in range(3):
        assert dscale[i] is None or dshift[i] is None or dss[i] == dhs[i]
        dss[i] = dss[i] or dhs[i]
    L.check(_lib.lx_ln_modulate_bwd(x.data_ptr(), dxn.data_ptr(), _ptr(dres), dx.data_ptr(), x.stride(0), x.shape[0],
                                    x.shape[1], tile_meta.data_ptr(), sp, ss, dsp, dhp, dss, eps, _ptr(stats), _stream()),
            "lx_ln_modulate_bwd")


def _rms3(ws):
    p = _P3()
    for i, w in enumerate(ws):
        if w is not None:
            p[i] = w.data_ptr()
    return p


def qkv_post_fwd(pre, heads, tile_meta, q, k, v, rms_q, rms_k, rope, eps=1e-6):
    L.check(_lib.lx_qkv_post_fwd(pre.data_ptr(), pre.stride(0), pre.shape[0], heads, tile_meta.data_ptr(), q.data_ptr(),
                                 k.data_ptr(), v.data_ptr(), q.shape[2], _rms3(rms_q), _rms3(rms_k), _ptr(rope), eps,
                                 _stream()), "lx_qkv_post_fwd")


def qkv_post_bwd(pre, dq, dk, dv, dpre, heads, tile_meta, rms_q, rms_k, rope, eps=1e-6):
    """dq: bf16, or the fp32 accumulation buffer of attention_bwd (read directly)."""
    fn = _lib.lx_qkv_post_bwd_f32dq if dq.dtype == torch.float32 else _lib.lx_qkv_post_bwd
    L.check(fn(pre.data_ptr(), pre.stride(0), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dpre.data_ptr(),
                                 dpre.stride(0), pre.shape[0], heads, tile_meta.data_ptr(), dq.shape[2], _rms3(rms_q),
                                 _rms3(rms_k), _ptr(rope), eps, _stream()), "lx_qkv_post_bwd")


def rows_to_heads(rows, heads, tile_meta, out):
    L.check(_lib.lx_rows_to_heads(rows.data_ptr(), rows.stride(0), out.data_ptr(), rows.shape[0], heads, tile_meta.data_ptr(),
                                  out.shape[2], _stream()), "lx_rows_to_heads")


def lora_grad(x, dy, A, Bw, dA, dB, scaling, workspace):
    M, K = x.shape
    N, r = Bw.shape
    assert dy.shape == (M, N) and A.shape == (r, K) and dA.shape == A.shape and dB.shape == Bw.shape
    assert A.is_contiguous() and Bw.is_contiguous() and dA.is_contiguous() and dB.is_contiguous()
    assert A.dtype == Bw.dtype == dA.dtype == dB.dtype == torch.float32 and workspace.numel() >= 2 * M * r
    L.check(_lib.lx_lora_grad(x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), A.data_ptr(), Bw.data_ptr(),
                              dA.data_ptr(), dB.data_ptr(), M, K, N, r, float(scaling), workspace.data_ptr(), _stream()),
            "lx_lora_grad")


class LoraStack(C.Structure):
    """lx_lora_stack_t (include/loongx_b200.h)."""
    _fields_ = [("groups", C.c_int32), ("r", C.c_int32), ("width", C.c_int32 * 4), ("A", c_void_p * 4), ("B", c_void_p * 4),
                ("dA", c_void_p * 4), ("dB", c_void_p * 4), ("scaling", C.c_float * 4)]


_lib.lx_lora_grad_stacked.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, C.POINTER(LoraStack), c_void_p,
                                      c_void_p]


def lora_grad_stackable(factors) -> bool:
    """can lx_lora_grad_stacked take these sub-Linears (sharing x, adjacent outputs) in one pass?"""
    r = factors[0].A.shape[0]
    return (r in (4, 8, 16) and len(factors) * r <= 16 and len(factors) <= 4 and factors[0].A.shape[1] % 8 == 0 and
            all(f.A.shape[0] == r and f.rows % 256 == 0 and f.A.data_ptr() % 16 == 0 and f.B.data_ptr() % 16 == 0 and
                f.dA.data_ptr() % 16 == 0 and f.dB.data_ptr() % 16 == 0 for f in factors))


def lora_grad_stacked(x, dy, factors, workspace):
    """dA / dB of `factors` (LoraFactor list): x [M, K] shared, outputs = adjacent column blocks of dy [M, sum widths]."""
    M, K = x.shape
    st = LoraStack()
    st.groups, st.r = len(factors), factors[0].A.shape[0]
    assert dy.shape == (M, sum(f.rows for f in factors)) and workspace.numel() >= 2 * M * st.groups * st.r
    assert x.stride(1) == 1 and dy.stride(1) == 1 and x.stride(0) % 8 == 0 and dy.stride(0) % 8 == 0
    for g, f in enumerate(factors):
        assert f.A.shape == (st.r, K) and f.B.shape == (f.rows, st.r) and f.A.is_contiguous() and f.B.is_contiguous()
        st.width[g], st.scaling[g] = f.rows, float(f.panel.scaling)
        st.A[g], st.B[g], st.dA[g], st.dB[g] = f.A.data_ptr(), f.B.data_ptr(), f.dA.data_ptr(), f.dB.data_ptr()
    L.check(_lib.lx_lora_grad_stacked(x.data_ptr(), x.stride(0), dy.data_ptr(), dy.stride(0), M, K, C.byref(st),
                                      workspace.data_ptr(), _stream()), "lx_lora_grad_stacked")


_lib.lx_lora_merge_t.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32,
                                 c_int32, C.c_float, c_void_p]


def lora_merge_t(W, A, Bw, out, outT, scaling):
    """out = bf16(W + s B A) and outT = out^T in one pass (outT: a [K, N] column view of the K-major panel)."""
    N, K = W.shape
    assert out.shape == (N, K) and outT.shape == (K, N) and outT.stride(1) == 1 and out.stride(1) == 1
    L.check(_lib.lx_lora_merge_t(W.data_ptr(), W.stride(0), A.data_ptr(), Bw.data_ptr(), out.data_ptr(), out.stride(0),
                                 outT.data_ptr(), outT.stride(0), N, K, A.shape[0], float(scaling), _stream()), "lx_lora_merge_t")


def lora_merge(W, A, Bw, out, scaling):
    N, K = W.shape
    L.check(_lib.lx_lora_merge(W.data_ptr(), W.stride(0), A.data_ptr(), Bw.data_ptr(), out.data_ptr(), out.stride(0), N, K,
                               A.shape[0], float(scaling), _stream()), "lx_lora_merge")


def transpose(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows, cols = x.shape
    if out is None:
        out = torch.empty((cols, rows), device=x.device, dtype=torch.bfloat16)
    L.check(_lib.lx_transpose_bf16(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, cols, _stream()),
            "lx_transpose_bf16")
    return out


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    L.check(_lib.lx_cast(x.data_ptr(), out.data_ptr(), x.numel(), 1, _stream()), "lx_cast")
    return out


def flow_noise_mix(x0, x1, t):
    xt = torch.empty_like(x0)
    L.check(_lib.lx_flow_noise_mix(x0.data_ptr(), x1.data_ptr(), t.data_ptr(), xt.data_ptr(), x0.shape[0],
                                   x0.numel() // x0.shape[0], _stream()), "lx_flow_noise_mix")
    return xt


def flow_mse_loss(pred, x0, x1, loss, dpred, grad_scale=1.0):
    L.check(_lib.lx_flow_mse_loss(pred.data_ptr(), x0.data_ptr(), x1.data_ptr(), loss.data_ptr(), _ptr(dpred), pred.numel(),
                                  float(grad_scale), _stream()), "lx_flow_mse_loss")


def sdpa_backward_library(q, k, v, d_out, n_cond: int, mask_mode: int, cross_bias: float):
    """LIBRARY CALL (interim): dQ, dK, dV of softmax(q k^T / sqrt(128) + mask) v through torch's
    scaled_dot_product_attention autograd, with the block masks / c_factor bias of block.py:106-128."""
    import torch.nn.functional as F

    S = q.shape[2]
    mask = None
    if n_cond > 0:
        if cross_bias != 0.0:
            mask = torch.zeros(S, S, device=q.device, dtype=q.dtype)
            mask[-n_cond:, :-n_cond] = cross_bias
            mask[:-n_cond, -n_cond:] = cross_bias
        elif mask_mode == 1:
            mask = torch.ones(S, S, device=q.device, dtype=torch.bool)
            mask[-n_cond:, :-n_cond] = False
            mask[:-n_cond, -n_cond:] = False
        elif mask_mode == 2:
            mask = torch.ones(S, S, device=q.device, dtype=torch.bool)
            mask[-n_cond:, :-n_cond] = False
    with torch.enable_grad():
        q_, k_, v_ = (t.detach().requires_grad_(True) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q_, k_, v_, attn_mask=mask, dropout_p=0.0, is_causal=False)
        return torch.autograd.grad(o, (q_, k_, v_), d_out)


# ------------------------------------------------------------------------------------------------------------------
# trainable LoRA factors
# ------------------------------------------------------------------------------------------------------------------
class _LoraShared:
    """Per weight set, per LoRA target: the nn.Parameter wrappers of the fp32 master factors (created ONCE, so that an
    optimizer built from them keeps training the same objects whatever trainers come and go) and the factor versions the
    merged panel was last built from."""

    __slots__ = ("A", "B", "merged")

    def __init__(self, A: torch.Tensor, Bw: torch.Tensor):
        self.A = torch.nn.Parameter(A, requires_grad=True)
        self.B = torch.nn.Parameter(Bw, requires_grad=True)
        self.merged = (self.A._version, self.B._version)


def lora_shared(weights: DitWeights) -> Dict[str, _LoraShared]:
    """name -> shared Parameter record of every LoRA target of `weights` (cached on the weights object)."""
    cache = weights.__dict__.setdefault("_lora_shared", {})
    for panel in weights.named.values():
        if not isinstance(panel, PackedLinear):
            continue
        for (name, _row0, _rows, A, Bw) in panel.lora:
            if name not in cache:
                cache[name] = _LoraShared(A, Bw)
    return cache


def set_lora_scale(weights: DitWeights, scale: float) -> None:
    """peft `scale_lora_layers` (transformer.py:73-83): every merged panel becomes W + scale * (alpha / r) B A.  The scale
    in force is ONE attribute of the weight set (`weights.lora_scale`): every later re-merge (optimizer step, load_lora)
    uses it, and the training step resets it to 1 before its forward."""
    scale = float(scale)
    if scale == getattr(weights, "lora_scale", 1.0):
        return
    weights.lora_scale = scale
    shared = lora_shared(weights)
    for panel in weights.named.values():
        if not isinstance(panel, PackedLinear):
            continue
        for (name, row0, rows, _A, _B) in panel.lora:
            LoraFactor(name, panel, row0, rows, shared[name]).remerge(scale)


def lora_parameters(weights: DitWeights) -> List[torch.nn.Parameter]:
    """The trainable parameters of the reference's optimizer (model.py:513-524, 541): every LoRA factor, in a fixed
    order.  Available before any training step and stable across batch geometries."""
    out: List[torch.nn.Parameter] = []
    for rec in lora_shared(weights).values():
        out += [rec.A, rec.B]
    return out


class LoraFactor:
    """One LoRA-targeted Linear: fp32 master factors living inside a PackedLinear panel.  dA / dB are views into the
    trainer's single flat gradient buffer (one NCCL all-reduce covers every factor); the Parameters are shared by every
    trainer of the weight set (`lora_shared`)."""

    def __init__(self, name: str, panel: PackedLinear, row0: int, rows: int, shared: _LoraShared):
        self.name, self.panel, self.row0, self.rows = name, panel, row0, rows
        self.shared = shared
        self.A, self.B = shared.A, shared.B
        self.dA: Optional[torch.Tensor] = None
        self.dB: Optional[torch.Tensor] = None

    def remerge(self, lora_scale: float = 1.0):
        """w_lora[rows] = bf16(W + lora_scale * s B A) (+ the transposed panel) after the factors changed."""
        p = self.panel
        w, wl = p.w[self.row0:self.row0 + self.rows], p.w_lora[self.row0:self.row0 + self.rows]
        wt = p.w_loraT[:, self.row0:self.row0 + self.rows] if p.w_loraT is not None else None
        if wt is not None and self.rows % 8 == 0 and w.shape[1] % 8 == 0 and (w.data_ptr() | wl.data_ptr() | wt.data_ptr()) % 16 == 0 \
                and w.stride(0) % 8 == 0 and wl.stride(0) % 8 == 0 and wt.stride(0) % 8 == 0:
            lora_merge_t(w, self.A.data, self.B.data, wl, wt, p.scaling * lora_scale)  # both panels in one pass over W
        else:
            lora_merge(w, self.A.data, self.B.data, wl, p.scaling * lora_scale)
            if wt is not None:
                transpose(wl, wt)
        self.shared.merged = (self.A._version, self.B._version)

    def stale(self) -> bool:
        return self.shared.merged != (self.A._version, self.B._version)


ALLREDUCE_LOG: list = []  # (start event, end event, bytes, rank-alignment start event) per gradient all-reduce when ALLREDUCE_TIMING is on (bench.py)
ALLREDUCE_TIMING = False


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """DDP gradient synchronisation of the reference's training harness (Lightning DDP): ONE all-reduce
    over the flat fp32 gradient bucket, divided by the world size.  NCCL over NVLink on the GPUs; gloo in the CPU test."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat
    timing = ALLREDUCE_TIMING and flat.is_cuda
    if timing:
        # measurement mode (bench.py): a one-element all-reduce first lines the ranks up on the device, so that the events
        # around the bucket's all-reduce time the exchange itself; the wait for the slowest rank's backward (clock spread
        # under the power cap) is logged separately instead of being charged to the collective
        es, e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        es.record()
        dist.all_reduce(torch.zeros(1, device=flat.device), op=dist.ReduceOp.SUM, group=group)
        e0.record()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.mul_(1.0 / dist.get_world_size(group))
    if timing:
        e1.record()
        ALLREDUCE_LOG.append((e0, e1, flat.numel() * flat.element_size(), es))
    return flat


def _hand_out(tr, flat, enc):
    """Slices of the (all-reduced) flat bucket, in the order the parameters were passed to the autograd node."""
    grads, o = [], 0
    for f in tr.factors.values():
        grads.append(flat[o:o + f.A.numel()].view_as(f.A))
        o += f.A.numel()
        grads.append(flat[o:o + f.B.numel()].view_as(f.B))
        o += f.B.numel()
    if enc is not None:
        grads += enc.parameter_grads(flat[tr.n_lora_grad:])
    return grads


class EncoderBackward:
    """The CS3 / DGF half of a training step: the conditioning context of this step, the encoder parameters and their
    gradient views inside the trainer's flat bucket (after the LoRA factors)."""

    def __init__(self, model, ctx, params, trainer):
        from . import cs3_bwd as CB

        self.CB, self.model, self.ctx, self.params = CB, model, ctx, list(params)
        assert trainer.grad_extra.numel() >= CB.grad_elements(self.params)
        self.views = CB.grad_views(self.params, trainer.grad_extra)

    def backward(self, d_prompt: torch.Tensor, d_pooled: torch.Tensor) -> None:
        self.CB.step_conditioning_backward(self.model, self.ctx, d_prompt.contiguous(), d_pooled.contiguous(), self.views)

    def parameter_grads(self, flat_extra: torch.Tensor):
        out = []
        for p, (o, n) in zip(self.params, self.CB.grad_layout(self.params)[0]):
            sl = flat_extra[o:o + n]
            out.append(torch.view_as_complex(sl.view(*p.shape, 2)) if p.is_complex() else sl.view_as(p))
        return out


class FlowStepFunction(torch.autograd.Function):
    """loss = step(batch) as an autograd node: `loss.backward()` (what Lightning calls on the reference's step output)
    runs the native backward (DiT, then - with `enc` - the CS3 / DGF conditioning), all-reduces the ONE flat gradient bucket
    across ranks and hands every LoRA factor and every encoder parameter its slice."""

    @staticmethod
    def forward(ctx, trainer, inputs, enc, *params):
        loss = trainer.forward(*inputs)
        ctx.trainer, ctx.enc = trainer, enc
        return loss.clone().reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        tr, enc = ctx.trainer, ctx.enc
        tr.zero_grad()
        tr.backward(float(grad_out))
        if enc is not None:
            enc.backward(tr.d_prompt, tr.d_pooled)
        allreduce_mean_(tr.grad_flat)
        # hand autograd a private copy: the trainer re-zeroes its bucket on the next backward, while `.grad` must keep
        # accumulating across micro-batches (accumulate_grad_batches: 4 in train/config/seed_512.yaml:12)
        flat = tr.grad_flat.clone()
        return (None, None, None, *_hand_out(tr, flat, enc))


class FlowStepMicroFunction(torch.autograd.Function):
    """A step over k micro-batches (gradient accumulation inside the step): every micro-batch runs forward + backward
    right away (its activations are dropped before the next one), the gradients accumulate in the trainer's flat bucket
    and `loss.backward()` only all-reduces / scales / hands them out.  loss = mean of the micro-batch losses.  The
    conditioning was computed for the whole batch: its backward runs once, on the concatenated input gradients."""

    @staticmethod
    def forward(ctx, trainer, chunks, enc, *params):
        k = len(chunks)
        trainer.zero_grad()
        total = torch.zeros((), device=trainer.loss.device, dtype=torch.float32)
        d_pe, d_po = [], []
        for inputs in chunks:
            total += trainer.forward(*inputs).reshape(())
            trainer.backward(1.0 / k)
            if enc is not None:
                d_pe.append(trainer.d_prompt.clone())
                d_po.append(trainer.d_pooled.clone())
        if enc is not None:
            enc.backward(torch.cat(d_pe, 0), torch.cat(d_po, 0))
        ctx.trainer, ctx.enc = trainer, enc
        return total / k

    @staticmethod
    def backward(ctx, grad_out):
        tr, enc = ctx.trainer, ctx.enc
        allreduce_mean_(tr.grad_flat)
        flat = tr.grad_flat * grad_out.to(tr.grad_flat.dtype)
        return (None, None, None, *_hand_out(tr, flat, enc))


class DitTrainer:
    """Native forward + backward of the rectified-flow objective for one batch geometry."""

    def __init__(self, weights: DitWeights, B: int, n_txt: int, n_img: int, n_cond: int, model_config: Optional[dict] = None,
                 attn_bwd: str = "native", recompute: Optional[bool] = None, input_grads: bool = False,
                 extra_grad_elems: int = 0):
        """recompute: True = checkpoint the residual stream per block and rebuild the block's intermediates in the
        backward (the reference's gradient_checkpointing, train/config/seed_512.yaml:16); False = keep every block's
        intermediates from the forward (15.5 GB per 512x512 sample for FLUX.1-dev: no second forward, ~1.3x faster);
        None = keep them when they fit in the free HBM, else recompute.  Same gradients either way.
        input_grads: also produce d loss / d prompt_embeds and d loss / d pooled_projections (`self.d_prompt`, `self.d_pooled`
        after backward()): what flows back into the CS3 / DGF conditioning of OminiModel.step (model.py:656-701).
        extra_grad_elems: fp32 elements appended to the flat gradient bucket after the LoRA factors (`self.grad_extra`), so
        that ONE all-reduce covers the encoder gradients as well."""
        model_config = model_config or {}
        assert attn_bwd in ("native", "library")
        self.attn_bwd = attn_bwd
        self.latent_lora = bool(model_config.get("latent_lora", False))
        self.input_grads = bool(input_grads)
        # block.py:233-234: the gated attention output of the condition stream is also added to the image stream
        self.add_cond_attn = bool(model_config.get("add_cond_attn", False))
        if self.add_cond_attn and n_cond != n_img:
            raise ValueError("model_config.add_cond_attn needs condition and image streams of the same length (block.py:234)")
        if n_cond <= 0:
            raise NotImplementedError("the training step needs a condition stream (the LoRA lives on the condition branch)")
        self.w, self.cfg = weights, weights.cfg
        self.plan = DitPlan(weights, B, n_txt, n_img, n_cond, T=1, model_config=model_config)
        # ragged stream lengths: the plan pads every stream to 128-token tiles; all row counts below are PADDED, the loss
        # and dpred cover the valid image tokens only (padding rows carry zero gradients, padding keys are masked)
        self.ni_valid, self.nc_valid = n_img, n_cond
        n_txt, n_img, n_cond = self.plan.ntp, self.plan.nip, self.plan.ncp
        self.pads = (self.plan.ntp - self.plan.nt, self.plan.nip - self.plan.ni, self.plan.ncp - self.plan.nc)
        self.B, self.nt, self.ni, self.nc = B, n_txt, n_img, n_cond
        cfg = self.cfg
        self.D, self.H = cfg.inner_dim, cfg.num_attention_heads
        D, dev = self.D, weights.device
        S = n_txt + n_img + n_cond
        self.S, self.R = S, B * S
        self.Rt, self.Ri, self.Rc = B * n_txt, B * n_img, B * n_cond
        R = self.R
        bf = dict(device=dev, dtype=torch.bfloat16)
        z = lambda *s: torch.zeros(s, **bf)  # noqa: E731
        nl_, ns_ = cfg.num_layers, cfg.num_single_layers
        if recompute is None:
            recompute = not self.activations_fit(weights, B, S)
        self.recompute = bool(recompute)
        pb = self.plan.buf
        # shared workspace: everything in recompute mode; in store mode only the pieces no backward kernel reads
        self.a = dict(XN=pb["XN"], XN2=z(R, D), Q=pb["Q"], K=pb["K"], V=pb["V"],
                      lse=torch.zeros((B, self.H, S), device=dev, dtype=torch.float32))
        if self.recompute:
            self.a.update(QM=z(R, 7 * D), Cat=z(R, 5 * D), Y1=z(R, D), Hid=z(R, 4 * D), Y2=z(R, D))
            self.slabs = None
        else:
            self.slabs = [self._slab(i < nl_, z, dev) for i in range(nl_ + ns_)]
        self.g = dict(dX=z(R, D), dX1=z(R, D), dY=z(R, D), dXN=z(R, D), dBig=z(R, 7 * D), dCat=z(R, 5 * D),
                      dOh=z(B, self.H, S, 128), dKh=z(B, self.H, S, 128), dVh=z(B, self.H, S, 128))
        self.delta = torch.zeros((B, self.H, S), device=dev, dtype=torch.float32)
        self.dq32 = torch.zeros((B, self.H, S, 128), device=dev, dtype=torch.float32)
        self.stats = torch.zeros((R, 2), device=dev, dtype=torch.float32)
        self.lora_ws = torch.zeros((2 * max(R, B) * max(cfg.lora_rank, 16),), device=dev, dtype=torch.float32)
        self.stack_lora = os.environ.get("LX_LORA_STACK", "1") != "0"  # A/B knob: per-factor kernels instead
        # q/k/v post-processing (RMSNorm, RoPE, head scatter) in the GEMM epilogue, as at inference, with the pre-norm projection
        # kept (qkv_pre): measured no faster than the separate row kernel (the 256-column tile the epilogue needs costs what
        # the kernel saves: 450 vs 359 + 83 us per double block at B = 4), so off unless LX_FUSE_QKV=1
        self.fuse_qkv = os.environ.get("LX_FUSE_QKV", "0") == "1" and self.H % 2 == 0
        self.fuse_gate = os.environ.get("LX_FUSE_GATE", "1") != "0"  # A/B knob: separate gate + residual kernel
        self.fuse_gelu = os.environ.get("LX_FUSE_GELU", "1") != "0"  # A/B knob: separate GELU forward / backward kernels
        self.ckpt = torch.zeros((cfg.num_layers + cfg.num_single_layers, R, D), **bf)
        self.ckpt_mid = torch.zeros((max(cfg.num_layers, 1), R, D), **bf)  # residual stream after the attention branch
        self.dmod_dbl = torch.zeros((B, max(cfg.num_layers, 1) * 6 * D), device=dev, dtype=torch.float32)
        self.dmod_sgl = torch.zeros((B, max(cfg.num_single_layers, 1) * 3 * D), device=dev, dtype=torch.float32)
        if self.latent_lora or self.input_grads:  # modulation gradients of the image stream (double) / the shared text+image vector (single)
            self.dmod_dbl_img = torch.zeros_like(self.dmod_dbl)
            self.dmod_sgl_ti = torch.zeros_like(self.dmod_sgl)
        if self.input_grads:  # ... and of the text stream / the final AdaLayerNormContinuous: temb depends on `pooled`
            self.dmod_dbl_txt = torch.zeros_like(self.dmod_dbl)
            self.dmod_out = torch.zeros((B, 2 * D), device=dev, dtype=torch.float32)
            self.d_prompt: Optional[torch.Tensor] = None
            self.d_pooled: Optional[torch.Tensor] = None
        self.loss = torch.zeros((1,), device=dev, dtype=torch.float32)
        self.factors: Dict[str, LoraFactor] = {}
        shared = lora_shared(weights)
        for key, panel in weights.named.items():
            if not isinstance(panel, PackedLinear):
                continue
            for (name, row0, rows, A, Bw) in panel.lora:
                self.factors[name] = LoraFactor(name, panel, row0, rows, shared[name])
        n_grad = sum(f.A.numel() + f.B.numel() for f in self.factors.values())
        self.grad_flat = torch.zeros((n_grad + int(extra_grad_elems),), device=dev, dtype=torch.float32)
        self.n_lora_grad = n_grad
        self.grad_extra = self.grad_flat[n_grad:]
        o = 0
        for f in self.factors.values():
            f.dA = self.grad_flat[o:o + f.A.numel()].view_as(f.A)
            o += f.A.numel()
            f.dB = self.grad_flat[o:o + f.B.numel()].view_as(f.B)
            o += f.B.numel()
        self._ensure_transposed()
        self._saved = None

    # -- weights -----------------------------------------------------------------------------------------------------
    def _ensure_transposed(self):
        for key, p in self.w.named.items():
            if not isinstance(p, PackedLinear) or not (key.startswith("double.") or key.startswith("single.") or key == "proj_out" or
                                                       (key == "context_embedder" and self.input_grads)):
                continue
            if p.wT is None:
                p.wT = transpose(p.w)
            if p.w_lora is not None and p.w_loraT is None:
                p.w_loraT = transpose(p.w_lora)

    def parameters(self) -> List[torch.nn.Parameter]:
        out = []
        for f in self.factors.values():
            out += [f.A, f.B]
        return out

    def named_parameters(self):
        for n, f in self.factors.items():
            yield n + ".lora_A.weight", f.A
            yield n + ".lora_B.weight", f.B

    def remerge(self):
        """Call after an optimizer step: rebuild every merged panel from the updated factors."""
        for f in self.factors.values():
            f.remerge(getattr(self.w, "lora_scale", 1.0))

    def remerge_if_stale(self):
        """Before a training forward: panels at LoRA scale 1 (a generate(..., joint_attention_kwargs={"scale": s}) may
        have left them at s) and rebuilt from the current factors."""
        set_lora_scale(self.w, 1.0)
        for f in self.factors.values():
            if f.stale():
                f.remerge(1.0)

    def zero_grad(self):
        self.grad_flat.zero_()

    def step_loss_micro(self, chunks, enc: Optional["EncoderBackward"] = None) -> torch.Tensor:
        """`chunks`: list of input tuples (each of this trainer's batch size) -> differentiable mean loss."""
        self.remerge_if_stale()
        return FlowStepMicroFunction.apply(self, chunks, enc, *self.parameters(), *(enc.params if enc is not None else ()))

    def step_loss(self, *inputs, enc: Optional["EncoderBackward"] = None) -> torch.Tensor:
        """Differentiable loss (0-dim fp32): forward now, native backward when autograd reaches it."""
        self.remerge_if_stale()
        return FlowStepFunction.apply(self, inputs, enc, *self.parameters(), *(enc.params if enc is not None else ()))

    # -- per-block helpers ---------------------------------------------------------------------------------------------
    def _mods_double(self, i):
        """six [B, D] chunk views per stream of block i: [shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp]."""
        D, b = self.D, self.plan.buf
        sl = lambda t, c: t[:, (i * 6 + c) * D:(i * 6 + c + 1) * D]  # noqa: E731
        return [[sl(b["mod_txt"], c), sl(b["mod_img"], c), sl(b["mod_cond_img"], c)] for c in range(6)]

    def _mods_single(self, i):
        D, b = self.D, self.plan.buf
        sl = lambda t, c: t[:, (i * 3 + c) * D:(i * 3 + c + 1) * D]  # noqa: E731
        return [[sl(b["mod_single"], c), sl(b["mod_single"], c), sl(b["mod_cond_single"], c)] for c in range(3)]

    def _groups(self, main: PackedLinear, ctx: Optional[PackedLinear], transposed: bool):
        """row groups of a GEMM over all R rows -> (W0, bias0, extra groups)."""
        pick = (lambda p, lora: (p.w_loraT if lora and p.w_loraT is not None else p.wT)) if transposed else \
               (lambda p, lora: (p.w_lora if lora and p.w_lora is not None else p.w))
        bias = (lambda p: None) if transposed else (lambda p: p.bias)
        ll = self.latent_lora  # model_config.latent_lora: the adapters stay active on the image (single: text+image) rows
        if ctx is not None:  # double block: [txt | img | cond]
            return pick(ctx, False), bias(ctx), [(pick(main, ll), bias(main), self.Rt),
                                                 (pick(main, True), bias(main), self.Rt + self.Ri)]
        return pick(main, ll), bias(main), [(pick(main, True), bias(main), self.Rt + self.Ri)]  # single: [txt+img | cond]

    def _gemm(self, A, main, ctx, out, transposed=False, mode=L.EPI_BIAS, **kw):
        W0, b0, extra = self._groups(main, ctx, transposed)
        ops.gemm(A, W0, b0, out, mode, groups=extra, **kw)

    def _lora_grads(self, names, x_rows, dy_rows, col0=0):
        """accumulate dA / dB of consecutive sub-Linears `names` whose outputs are adjacent column blocks of dy_rows."""
        fs = [self.factors[n] for n in names]
        c = col0
        if self.stack_lora:
            per = max(1, 16 // max(fs[0].A.shape[0], 1))  # sub-Linears per pass: groups * rank <= 16
            while fs and lora_grad_stackable(fs[:per]) and (x_rows.data_ptr() | dy_rows[:, c:].data_ptr()) % 16 == 0:
                part, fs = fs[:per], fs[per:]
                width = sum(f.rows for f in part)
                lora_grad_stacked(x_rows, dy_rows[:, c:c + width], part, self.lora_ws)
                c += width
        for f in fs:  # ranks / widths the stacked kernels do not take, tiny row counts
            lora_grad(x_rows, dy_rows[:, c:c + f.rows], f.A.data, f.B.data, f.dA, f.dB, f.panel.scaling, self.lora_ws)
            c += f.rows

    @staticmethod
    def activation_bytes(cfg, B: int, S: int) -> int:
        """HBM needed to keep every block's backward operands (store mode): 18 / 17 [R, D] bf16 buffers per double /
        single block (see _slab) + the log-sum-exp rows."""
        R, D, H = B * S, cfg.inner_dim, cfg.num_attention_heads
        return (cfg.num_layers * 18 + cfg.num_single_layers * 17) * R * D * 2 + (cfg.num_layers + cfg.num_single_layers) * B * H * S * 4

    @staticmethod
    def activations_fit(weights: DitWeights, B: int, S: int) -> bool:
        free, _ = torch.cuda.mem_get_info(weights.device)
        already = any(isinstance(p, PackedLinear) and p.wT is not None for p in weights.named.values())
        w_bytes = 0 if already else weights.param_bytes()  # transposed panels (built once) ~ as large as the block weights
        return DitTrainer.activation_bytes(weights.cfg, B, S) + w_bytes + (10 << 30) <= free

    def _slab(self, double: bool, z, dev):
        """per-block activations the backward reads (store mode): modulated input, pre-activation projections, attention
        operands + log-sum-exp, attention output / single-block concat, pre-gate projection outputs, FF hidden."""
        B, H, S, R, D = self.B, self.H, self.S, self.R, self.D
        a = dict(XN=z(R, D), QM=z(R, 7 * D), Cat=z(R, D if double else 5 * D), Y1=z(R, D), Q=z(B, H, S, 128),
                 K=z(B, H, S, 128), V=z(B, H, S, 128), lse=torch.zeros((B, H, S), device=dev, dtype=torch.float32),
                 XN2=self.a["XN2"])
        if double:
            a.update(Hid=z(R, 4 * D), Y2=z(R, D))
        return a

    def _act(self, blk: int):
        """activation set of block `blk` (double blocks first): its own slab, or the shared recompute workspace."""
        return self.a if self.slabs is None else self.slabs[blk]

    def _attention(self, out, a):
        """out: [R, ld] view; head h lands in columns [128h, 128h+128).  Also records the log-sum-exp rows."""
        b, p = self.plan.buf, self.plan.plan
        ops.attention(a["Q"], a["K"], a["V"], out, b["out_row_base"], n_cond=self.nc, mask_mode=p.mask_mode,
                      cross_bias=p.cross_bias, lse=a["lse"], pads=self.pads, n_txt=self.nt)

    def _attention_bwd(self, d_rows, o_rows, a):
        """d_rows / o_rows: [R, >= D] views holding dO / O in their first D columns -> dq, dk, dv head-major (bf16)."""
        b, p, g = self.plan.buf, self.plan.plan, self.g
        if self.attn_bwd == "library":  # A/B check only: torch SDPA autograd
            rows_to_heads(d_rows, self.H, b["tile_meta"], g["dOh"])
            return sdpa_backward_library(a["Q"], a["K"], a["V"], g["dOh"], self.nc, p.mask_mode, p.cross_bias)
        ops.attention_bwd_prep(d_rows, o_rows, self.H, b["tile_meta"], g["dOh"], self.delta)
        self.dq32.zero_()
        ops.attention_bwd(a["Q"], a["K"], a["V"], g["dOh"], a["lse"], self.delta, self.dq32, g["dKh"], g["dVh"],
                          n_cond=self.nc, mask_mode=p.mask_mode, cross_bias=p.cross_bias, pads=self.pads, n_txt=self.nt)
        return self.dq32, g["dKh"], g["dVh"]  # (qkv_post_bwd reads the fp32 dQ directly)

    # -- blocks -------------------------------------------------------------------------------------------------------
    def _gemm_cond(self, A, main: PackedLinear, out, ctx: Optional[PackedLinear] = None):
        """the condition rows only, against the LoRA-merged panel (recompute of a projection whose output the backward
        needs on the condition stream alone: the gate gradient feeds the LoRA of the AdaLN linear).  With latent_lora the
        image (single blocks: text + image) rows carry LoRA as well, so everything is rebuilt."""
        if self.latent_lora:
            return self._gemm(A, main, ctx, out)
        c0 = self.Rt + self.Ri
        ops.gemm(A[c0:], main.w_lora if main.w_lora is not None else main.w, main.bias, out[c0:], L.EPI_BIAS)

    def _double_fwd(self, i, recompute: bool = False):
        """recompute=True (backward): the residual stream after the attention branch comes from its checkpoint, the
        pre-gate projection outputs are rebuilt for the condition rows only and the block output is not formed."""
        b, a, D, W = self.plan.buf, self._act(i), self.D, self.w.named
        X, tm = b["X"], b["tile_meta"]
        m = self._mods_double(i)
        pre = a["QM"][:, :3 * D]
        ln_modulate(X, a["XN"], tm, m[0], m[1])
        nq, nk, naq, nak = (W[f"double.{i}.{n}"] for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"))
        if self.fuse_qkv:  # the inference epilogue (RMSNorm, RoPE, head scatter) + the pre-norm projection kept for the backward
            self._gemm(a["XN"], W[f"double.{i}.qkv"], W[f"double.{i}.qkv_ctx"], None, mode=L.EPI_QKV, tile_meta=tm,
                       qkv=(a["Q"], a["K"], a["V"]), rms_q=[naq, nq, nq], rms_k=[nak, nk, nk], rope=b["rope"], qkv_pre=pre)
        else:
            self._gemm(a["XN"], W[f"double.{i}.qkv"], W[f"double.{i}.qkv_ctx"], pre)
            qkv_post_fwd(pre, self.H, tm, a["Q"], a["K"], a["V"], [naq, nq, nq], [nak, nk, nk], b["rope"])
        O = a["Cat"][:, :D]
        self._attention(O, a)
        if recompute:
            self._gemm_cond(O, W[f"double.{i}.out"], a["Y1"], W[f"double.{i}.out_ctx"])
            x1 = self.ckpt_mid[i]
        else:
            x1 = self.ckpt_mid[i]  # written in place: the checkpoint IS the forward's buffer
            if self.fuse_gate:  # x1 = X + gate * y from the GEMM epilogue, which also keeps y for the gate gradient
                self._gemm(O, W[f"double.{i}.out"], W[f"double.{i}.out_ctx"], x1, mode=L.EPI_GATE_RESIDUAL, residual=X,
                           gate=m[2], tile_meta=tm, out2=(a["Y1"], 0))
            else:
                self._gemm(O, W[f"double.{i}.out"], W[f"double.{i}.out_ctx"], a["Y1"])
                gate_residual_fwd(X, a["Y1"], x1, tm, m[2])
            if self.add_cond_attn:  # block.py:233-234: hidden_states += cond_gate_msa * cond_attn_output (row for row)
                ri, rc = self.Rt, self.Rt + self.Ri
                gate_residual_fwd(x1[ri:rc], a["Y1"][rc:], x1[ri:rc], tm[rc // 128:], m[2])
        ln_modulate(x1, a["XN2"], tm, m[3], m[4])
        pre_ff = a["QM"][:, 3 * D:7 * D]
        if self.fuse_gelu:  # one launch writes the GELU and, where the pre-activation would go, gelu' for the backward
            self._gemm(a["XN2"], W[f"double.{i}.ff_up"], W[f"double.{i}.ff_ctx_up"], pre_ff, mode=L.EPI_BIAS_GELU_DUAL,
                       out2=(a["Hid"], 0))
        else:
            self._gemm(a["XN2"], W[f"double.{i}.ff_up"], W[f"double.{i}.ff_ctx_up"], pre_ff)
            gelu_fwd(pre_ff, a["Hid"])
        if recompute:
            self._gemm_cond(a["Hid"], W[f"double.{i}.ff_down"], a["Y2"], W[f"double.{i}.ff_ctx_down"])
        else:
            if self.fuse_gate:
                self._gemm(a["Hid"], W[f"double.{i}.ff_down"], W[f"double.{i}.ff_ctx_down"], X, mode=L.EPI_GATE_RESIDUAL,
                           residual=x1, gate=m[5], tile_meta=tm, out2=(a["Y2"], 0))
            else:
                self._gemm(a["Hid"], W[f"double.{i}.ff_down"], W[f"double.{i}.ff_ctx_down"], a["Y2"])
                gate_residual_fwd(x1, a["Y2"], X, tm, m[5])

    def _double_bwd(self, i, x_in):
        """gradient wrt the block output is in g['dX']; leaves the gradient wrt the block input there."""
        b, a, g, D, W = self.plan.buf, self._act(i), self.g, self.D, self.w.named
        tm = b["tile_meta"]
        m = self._mods_double(i)
        c0 = self.Rt if self.latent_lora else self.Rt + self.Ri  # first row whose Linear carries LoRA
        mg = self.latent_lora or self.input_grads
        dm = lambda c: [self.dmod_dbl_txt[:, (i * 6 + c) * D:(i * 6 + c + 1) * D] if self.input_grads else None,  # noqa: E731
                        self.dmod_dbl_img[:, (i * 6 + c) * D:(i * 6 + c + 1) * D] if mg else None,
                        self.dmod_dbl[:, (i * 6 + c) * D:(i * 6 + c + 1) * D]]
        pfx = f"transformer_blocks.{i}."
        pre, pre_ff = a["QM"][:, :3 * D], a["QM"][:, 3 * D:7 * D]
        O = a["Cat"][:, :D]
        # feed-forward branch
        gate_bwd(g["dX"], a["Y2"], g["dY"], tm, m[5], dm(5))
        self._lora_grads([pfx + "ff.net.2"], a["Hid"][c0:], g["dY"][c0:])
        d_hid = g["dBig"][:, :4 * D]
        if self.fuse_gelu:  # d_hid = (dY W_down) * gelu' in the GEMM epilogue (pre_ff holds gelu', see _double_fwd)
            self._gemm(g["dY"], W[f"double.{i}.ff_down"], W[f"double.{i}.ff_ctx_down"], d_hid, transposed=True,
                       mode=L.EPI_MUL_AUX, residual=pre_ff)
        else:
            self._gemm(g["dY"], W[f"double.{i}.ff_down"], W[f"double.{i}.ff_ctx_down"], d_hid, transposed=True)
            gelu_bwd(pre_ff, d_hid, d_hid)
        self._gemm(d_hid, W[f"double.{i}.ff_up"], W[f"double.{i}.ff_ctx_up"], g["dXN"], transposed=True)
        ln_modulate_bwd(self.ckpt_mid[i], g["dXN"], g["dX"], g["dX1"], tm, m[4], dm(4), dm(3), self.stats)
        # attention branch
        if self.add_cond_attn:
            # the condition's gated projection also feeds the image stream (block.py:233-234): its upstream gradient is
            # d x1[cond] + d x1[img]; the residual path of the condition rows keeps d x1[cond] alone (dXN is free here)
            ri, rc = self.Rt, self.Rt + self.Ri
            d_c = g["dXN"][rc:]
            d_c.copy_(g["dX1"][rc:])
            add_rows(d_c, g["dX1"][ri:rc])
            gate_bwd(g["dX1"][:rc], a["Y1"][:rc], g["dY"][:rc], tm[:rc // 128], m[2], dm(2))
            gate_bwd(d_c, a["Y1"][rc:], g["dY"][rc:], tm[rc // 128:], m[2], dm(2))
        else:
            gate_bwd(g["dX1"], a["Y1"], g["dY"], tm, m[2], dm(2))
        self._lora_grads([pfx + "attn.to_out.0"], O[c0:], g["dY"][c0:])
        d_o = g["dCat"][:, :D]
        self._gemm(g["dY"], W[f"double.{i}.out"], W[f"double.{i}.out_ctx"], d_o, transposed=True)
        dq, dk, dv = self._attention_bwd(d_o, O, a)
        nq, nk, naq, nak = (W[f"double.{i}.{n}"] for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"))
        d_pre = g["dBig"][:, :3 * D]
        qkv_post_bwd(pre, dq, dk, dv, d_pre, self.H, tm, [naq, nq, nq], [nak, nk, nk], b["rope"])
        self._lora_grads([pfx + "attn.to_q", pfx + "attn.to_k", pfx + "attn.to_v"], a["XN"][c0:], d_pre[c0:])
        self._gemm(d_pre, W[f"double.{i}.qkv"], W[f"double.{i}.qkv_ctx"], g["dXN"], transposed=True)
        ln_modulate_bwd(x_in, g["dXN"], g["dX1"], g["dX"], tm, m[1], dm(1), dm(0), self.stats)

    def _single_fwd(self, i, recompute: bool = False):
        b, a, D, W = self.plan.buf, self._act(self.cfg.num_layers + i), self.D, self.w.named
        X, tm = b["X"], b["tile_meta"]
        m = self._mods_single(i)
        ln_modulate(X, a["XN"], tm, m[0], m[1])
        fuse = self.fuse_gelu and D % 256 == 0  # (the same condition as _single_bwd: QM[:, 3D:] holds gelu', not the pre-activation)
        nq, nk = W[f"single.{i}.norm_q"], W[f"single.{i}.norm_k"]
        # columns [0, 3D) = q | k | v, [3D, 7D) = proj_mlp (fused: its GELU into the concat buffer, gelu' for the backward into QM)
        seg1 = dict(n_split=3 * D, seg1=(L.EPI_BIAS_GELU_DUAL, a["QM"], 3 * D), out2=(a["Cat"], D)) if fuse else \
            dict(n_split=3 * D, seg1=(L.EPI_BIAS, a["QM"], 3 * D))
        if self.fuse_qkv:
            self._gemm(a["XN"], W[f"single.{i}.qkv_mlp"], None, None, mode=L.EPI_QKV, tile_meta=tm, qkv=(a["Q"], a["K"], a["V"]),
                       rms_q=[nq, nq, nq], rms_k=[nk, nk, nk], rope=b["rope"], qkv_pre=a["QM"], **seg1)
        else:
            if fuse:
                self._gemm(a["XN"], W[f"single.{i}.qkv_mlp"], None, a["QM"], **seg1)
            else:
                self._gemm(a["XN"], W[f"single.{i}.qkv_mlp"], None, a["QM"])
            qkv_post_fwd(a["QM"], self.H, tm, a["Q"], a["K"], a["V"], [nq, nq, nq], [nk, nk, nk], b["rope"])
        if not fuse:
            gelu_fwd(a["QM"][:, 3 * D:], a["Cat"][:, D:])
        self._attention(a["Cat"], a)
        if recompute:
            self._gemm_cond(a["Cat"], W[f"single.{i}.proj_out"], a["Y1"])
        else:
            if self.fuse_gate:
                self._gemm(a["Cat"], W[f"single.{i}.proj_out"], None, X, mode=L.EPI_GATE_RESIDUAL, residual=X, gate=m[2],
                           tile_meta=tm, out2=(a["Y1"], 0))
            else:
                self._gemm(a["Cat"], W[f"single.{i}.proj_out"], None, a["Y1"])
                gate_residual_fwd(X, a["Y1"], X, tm, m[2])

    def _single_bwd(self, i, x_in):
        b, a, g, D, W = self.plan.buf, self._act(self.cfg.num_layers + i), self.g, self.D, self.w.named
        tm = b["tile_meta"]
        m = self._mods_single(i)
        c0 = 0 if self.latent_lora else self.Rt + self.Ri
        dm = lambda c: ([self.dmod_sgl_ti[:, (i * 3 + c) * D:(i * 3 + c + 1) * D]] * 2  # noqa: E731
                        if (self.latent_lora or self.input_grads) else [None, None]) + \
            [self.dmod_sgl[:, (i * 3 + c) * D:(i * 3 + c + 1) * D]]
        pfx = f"single_transformer_blocks.{i}."
        gate_bwd(g["dX"], a["Y1"], g["dY"], tm, m[2], dm(2))
        self._lora_grads([pfx + "proj_out"], a["Cat"][c0:], g["dY"][c0:])
        if self.fuse_gelu and D % 256 == 0:  # columns [D, 5D) of dCat times gelu' (QM[:, 3D:]) = d(proj_mlp pre-activation)
            self._gemm(g["dY"], W[f"single.{i}.proj_out"], None, g["dCat"], transposed=True, n_split=D,
                       seg1=(L.EPI_MUL_AUX, g["dBig"], 3 * D), residual=a["QM"])
        else:
            self._gemm(g["dY"], W[f"single.{i}.proj_out"], None, g["dCat"], transposed=True)
            gelu_bwd(a["QM"][:, 3 * D:], g["dCat"][:, D:], g["dBig"][:, 3 * D:])
        dq, dk, dv = self._attention_bwd(g["dCat"], a["Cat"], a)
        nq, nk = W[f"single.{i}.norm_q"], W[f"single.{i}.norm_k"]
        qkv_post_bwd(a["QM"], dq, dk, dv, g["dBig"], self.H, tm, [nq, nq, nq], [nk, nk, nk], b["rope"])
        self._lora_grads([pfx + "attn.to_q", pfx + "attn.to_k", pfx + "attn.to_v", pfx + "proj_mlp"], a["XN"][c0:],
                         g["dBig"][c0:])
        self._gemm(g["dBig"], W[f"single.{i}.qkv_mlp"], None, g["dXN"], transposed=True)
        ln_modulate_bwd(x_in, g["dXN"], g["dX"], g["dX"], tm, m[1], dm(1), dm(0), self.stats)  # in place: row-local

    # -- public ---------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x0, x1, t, cond_latents, prompt_embeds, pooled, txt_ids, img_ids, cond_ids, guidance=1.0):
        """x0, x1: bf16 [B, n_img, C] packed latents / noise; t: fp32 [B] in (0, 1) -> loss (fp32 [1] tensor).
        model.py:590-594 (x_t), 705-723 (tranformer_forward), 726 (mse)."""
        B, cfg, b, a = self.B, self.cfg, self.plan.buf, self.a
        assert x0.dtype == torch.bfloat16 and x0.shape == (B, self.ni_valid, cfg.in_channels) and x0.is_contiguous()
        assert self.attn_bwd == "native" or not any(self.pads), "the library A/B path has no padding mask"
        t = t.to(device=x0.device, dtype=torch.float32).contiguous()
        pad_tok = self.plan._pad_tokens
        x0, x1 = pad_tok(x0, self.ni), pad_tok(x1.contiguous(), self.ni)  # zero rows: pred - (x1 - x0) is masked below
        cond_latents = pad_tok(cond_latents.to(torch.bfloat16).contiguous(), self.nc)
        xt = flow_noise_mix(x0, x1, t)
        self.plan.set_ids(txt_ids, img_ids, cond_ids)
        ts = [float(v) for v in t.tolist()]
        self.plan.prepare(prompt_embeds, pooled, cond_latents[:, :self.nc_valid], ts,
                          [guidance] * B if cfg.guidance_embeds else None, c_t=0.0)
        self.plan.embed(xt[:, :self.ni_valid])
        X = b["X"]
        nl, ns = cfg.num_layers, cfg.num_single_layers
        for i in range(nl):
            self.ckpt[i].copy_(X)
            self._double_fwd(i)
        for i in range(ns):
            self.ckpt[nl + i].copy_(X)
            self._single_fwd(i)
        # norm_out (scale first, then shift) + proj_out on the image rows (transformer.py:241-244)
        Xi, XNi = X[self.Rt:self.Rt + self.Ri], a["XN"][self.Rt:self.Rt + self.Ri]
        tm_i = b["tile_meta"][self.Rt // 128:]
        mo = b["mod_out"]
        sc, sh = mo[:, :self.D], mo[:, self.D:]
        ln_modulate(Xi, XNi, tm_i, [sh, sh, sh], [sc, sc, sc])
        pred = torch.empty((B, self.ni, cfg.in_channels), device=X.device, dtype=torch.bfloat16)
        po = self.w.named["proj_out"]
        ops.gemm(XNi, po.w, po.bias, pred.view(self.Ri, cfg.in_channels), L.EPI_BIAS)
        if any(self.pads):  # padding image rows: make the residual exactly zero so they add nothing to loss / dpred
            pred[:, self.ni_valid:].zero_()
        self._saved = dict(x0=x0, x1=x1, pred=pred, xt=xt, cond_latents=cond_latents, X_final=X.clone(),
                           pooled=pooled.to(torch.bfloat16).contiguous())
        self.loss.zero_()
        flow_mse_loss(pred, x0, x1, self.loss, None)
        if any(self.pads):  # mean over the valid elements (the kernel divided by the padded count)
            self.loss.mul_(self.ni / self.ni_valid)
        self.pred = pred[:, :self.ni_valid]
        return self.loss

    @torch.no_grad()
    def backward(self, grad_scale: float = 1.0):
        """Accumulates d loss / d (LoRA A, B) into LoraFactor.dA / dB (fp32)."""
        assert self._saved is not None, "call forward() first"
        s, b, a, g, cfg = self._saved, self.plan.buf, self.a, self.g, self.cfg
        X, tm = b["X"], b["tile_meta"]
        nl, ns = cfg.num_layers, cfg.num_single_layers
        self.dmod_dbl.zero_()
        self.dmod_sgl.zero_()
        if self.latent_lora or self.input_grads:
            self.dmod_dbl_img.zero_()
            self.dmod_sgl_ti.zero_()
        if self.input_grads:
            self.dmod_dbl_txt.zero_()
            self.dmod_out.zero_()
        dpred = torch.empty_like(s["pred"])
        scratch_loss = torch.zeros_like(self.loss)
        flow_mse_loss(s["pred"], s["x0"], s["x1"], scratch_loss, dpred, grad_scale * (self.ni / self.ni_valid))
        # proj_out / norm_out backward on the image rows; text and condition rows of the final stream get no gradient
        g["dX"].zero_()
        sl = slice(self.Rt, self.Rt + self.Ri)
        po = self.w.named["proj_out"]
        ops.gemm(dpred.view(self.Ri, cfg.in_channels), po.wT, None, g["dXN"][sl], L.EPI_BIAS)
        mo = b["mod_out"]
        sc = mo[:, :self.D]
        if self.input_grads:  # AdaLayerNormContinuous: scale first, then shift (transformer.py:243)
            dsc, dsh = self.dmod_out[:, :self.D], self.dmod_out[:, self.D:]
            ln_modulate_bwd(s["X_final"][sl], g["dXN"][sl], None, g["dX"][sl], tm[self.Rt // 128:], [sc, sc, sc],
                            [None, dsc, None], [None, dsh, None], self.stats)
        else:
            ln_modulate_bwd(s["X_final"][sl], g["dXN"][sl], None, g["dX"][sl], tm[self.Rt // 128:], [sc, sc, sc],
                            [None] * 3, [None] * 3, None)
        for i in reversed(range(ns)):
            if self.recompute:  # gradient checkpointing, transformer.py:184-206
                X.copy_(self.ckpt[nl + i])
                self._single_fwd(i, recompute=True)
            self._single_bwd(i, self.ckpt[nl + i])
        for i in reversed(range(nl)):
            if self.recompute:
                X.copy_(self.ckpt[i])
                self._double_fwd(i, recompute=True)
            self._double_bwd(i, self.ckpt[i])
        # x_embedder on the condition rows (transformer.py:93) (+ the image rows with latent_lora, transformer.py:91-92)
        c0 = self.Rt + self.Ri
        self._lora_grads(["x_embedder"], s["cond_latents"].view(self.Rc, cfg.in_channels), g["dX"][c0:])
        if self.latent_lora:
            self._lora_grads(["x_embedder"], s["xt"].view(self.Ri, cfg.in_channels), g["dX"][self.Rt:c0])
        # AdaLN linears: emb -> [B, 6D | 3D] per block; input silu(cond_temb) for the condition stream, silu(temb) for the
        # image (double) / text+image (single) rows when latent_lora keeps their adapters active
        silu_c, silu_t = b["silu_c"], b["silu_t"]
        if self.input_grads:
            self._input_grads(s)
        pairs = [(silu_c, cast_bf16(self.dmod_dbl), cast_bf16(self.dmod_sgl))]
        if self.latent_lora:
            pairs.append((silu_t, cast_bf16(self.dmod_dbl_img), cast_bf16(self.dmod_sgl_ti)))
        for x_in, dmd, dms in pairs:  # every block's AdaLN Linear reads the same x: stacks of four per pass
            self._lora_grads([f"transformer_blocks.{i}.norm1.linear" for i in range(nl)], x_in, dmd)
            self._lora_grads([f"single_transformer_blocks.{i}.norm.linear" for i in range(ns)], x_in, dms)

    def _input_grads(self, s) -> None:
        """d loss / d prompt_embeds (through context_embedder, transformer.py:115) and d loss / d pooled_projections
        (through time_text_embed's text branch into temb and cond_temb, then every AdaLN Linear: transformer.py:102-114,
        block.py:192-207, 238-253, 301-305).  fp32 results in self.d_prompt [B, n_txt, joint_dim], self.d_pooled [B, pooled_dim]."""
        from . import cs3_bwd as CB

        b, g, cfg, W = self.plan.buf, self.g, self.cfg, self.w.named
        B, D, dev = self.B, self.D, b["X"].device
        assert B <= 8, "input gradients: micro-batches of at most 8 samples (skinny kernels)"
        # text tokens: X0_txt = prompt_embeds W_ctx^T + b
        ce = W["context_embedder"]
        d_pe = torch.empty((self.Rt, cfg.joint_attention_dim), device=dev, dtype=torch.float32)
        ops.gemm(g["dX"][:self.Rt], ce.wT, None, d_pe, L.EPI_BIAS_F32)
        self.d_prompt = d_pe.view(B, self.nt, cfg.joint_attention_dim)[:, :self.plan.nt]
        # conditioning vectors: d silu(temb) (rows of step 0 .. B-1) and d silu(cond_temb)
        d_silu = torch.zeros((2 * B, D), device=dev, dtype=torch.float32)
        d_t, d_c = d_silu[:B], d_silu[B:]
        ll = self.latent_lora

        def skinny(dm, panel, lora, out):
            w = panel.w_lora if (lora and panel.w_lora is not None) else panel.w
            L.check(CB._lib.lx_skinny_xw_bf16(dm.data_ptr(), dm.stride(0), w.data_ptr(), w.stride(0), out.data_ptr(), out.stride(0), B,
                                              w.shape[0], w.shape[1], _stream()), "lx_skinny_xw_bf16")

        skinny(self.dmod_dbl_txt, W["mod_txt"], False, d_t)
        skinny(self.dmod_dbl_img, W["mod_img"], ll, d_t)
        skinny(self.dmod_sgl_ti, W["mod_single"], ll, d_t)
        skinny(self.dmod_out, W["norm_out"], False, d_t)
        skinny(self.dmod_dbl, W["mod_img"], True, d_c)
        skinny(self.dmod_sgl, W["mod_single"], True, d_c)
        # silu: temb = e_t + e_g + e_x (rows [0, B) of step 0; the cond rows follow at M - B)
        emb = b["emb_tmp"]  # [4, M, D]: hidden, e_t, e_g, e_x
        M = emb.shape[1]
        assert M == 2 * B, "training plans have T = 1"
        d_pre = torch.empty_like(d_silu)
        e_g = emb[2] if cfg.guidance_embeds else None
        L.check(CB._lib.lx_silu_bwd_sum(d_silu.data_ptr(), emb[1].data_ptr(), e_g.data_ptr() if e_g is not None else None,
                                        emb[3].data_ptr(), B, d_pre.data_ptr(), M, D, _stream()), "lx_silu_bwd_sum")
        d_ex = torch.zeros((B, D), device=dev, dtype=torch.float32)
        CB._axpy(d_ex, d_pre[:B].contiguous())
        CB._axpy(d_ex, d_pre[B:].contiguous())
        # text branch: e_x = W2 silu(W1 pooled + b1) + b2
        t1, t2 = W["text_1"], W["text_2"]
        d_h = torch.zeros((B, D), device=dev, dtype=torch.float32)
        L.check(CB._lib.lx_skinny_xw_bf16(d_ex.data_ptr(), D, t2.w.data_ptr(), t2.w.stride(0), d_h.data_ptr(), D, B, t2.w.shape[0],
                                          t2.w.shape[1], _stream()), "lx_skinny_xw_bf16")
        h_pre = torch.empty((B, D), device=dev, dtype=torch.float32)
        ops.gemm(s["pooled"], t1.w, t1.bias, h_pre, L.EPI_BIAS_F32)
        d_h1 = torch.empty_like(d_h)
        L.check(CB._lib.lx_silu_bwd_f32(d_h.data_ptr(), h_pre.data_ptr(), d_h1.data_ptr(), d_h.numel(), _stream()), "lx_silu_bwd_f32")
        d_po = torch.zeros((B, t1.w.shape[1]), device=dev, dtype=torch.float32)
        L.check(CB._lib.lx_skinny_xw_bf16(d_h1.data_ptr(), D, t1.w.data_ptr(), t1.w.stride(0), d_po.data_ptr(), d_po.stride(0), B,
                                          t1.w.shape[0], t1.w.shape[1], _stream()), "lx_skinny_xw_bf16")
        self.d_pooled = d_po

    def grads(self) -> Dict[str, torch.Tensor]:
        out = {}
        for n, f in self.factors.items():
            out[n + ".lora_A.weight"] = f.dA
            out[n + ".lora_B.weight"] = f.dB
        return out
