"""GPU parity tests of the tcgen05 joint-attention kernel against fp32 torch math (block.py:106-135 semantics).

Tolerance (stated): P is rounded to bf16 before the second GEMM and the output is rounded to bf16:
|err| <= 4e-3 + 2^-7 |ref|.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _mk(shape, scale, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


def _ref(q, k, v, n_cond, mask_mode, cross_bias):
    B, H, S, D = q.shape
    qf, kf, vf = q.float(), k.float(), v.float()
    logits = qf @ kf.transpose(-1, -2) / math.sqrt(D)
    if n_cond > 0:
        if cross_bias != 0.0:
            bias = torch.zeros(S, S, device=q.device)
            bias[-n_cond:, :-n_cond] = cross_bias
            bias[:-n_cond, -n_cond:] = cross_bias
            logits = logits + bias
        elif mask_mode in (1, 2):
            mask = torch.ones(S, S, dtype=torch.bool, device=q.device)
            mask[-n_cond:, :-n_cond] = False
            if mask_mode == 1:
                mask[:-n_cond, -n_cond:] = False
            logits = logits.masked_fill(~mask, float("-inf"))
    return torch.softmax(logits, dim=-1) @ vf  # [B,H,S,D]


def _run(B, H, nt, ni, nc, mask_mode=0, cross_bias=0.0, qscale=1.0, col_offset=0, extra_cols=0, atol=4e-3, rdiv=128):
    from loongx_b200 import ops

    S = nt + ni + nc
    q, k, v = _mk((B, H, S, 128), qscale, 1), _mk((B, H, S, 128), 1.0, 2), _mk((B, H, S, 128), 1.0, 3)
    ld = H * 128 + col_offset + extra_cols
    out = torch.full((B * S, ld), float("nan"), device="cuda", dtype=torch.bfloat16)
    orb = ops.make_out_row_base(B, nt, ni, nc, "cuda")
    ops.attention(q, k, v, out, orb, n_cond=nc, mask_mode=mask_mode, cross_bias=cross_bias, col_offset=col_offset)
    torch.cuda.synchronize()
    ref = _ref(q, k, v, nc, mask_mode, cross_bias)  # [B,H,S,D]
    # gather kernel output back to [B,H,S,D] through the same row map
    got = torch.empty_like(ref)
    rows = orb.cpu().tolist()
    for b in range(B):
        for t in range(S // 128):
            r0 = rows[b * (S // 128) + t]
            blk = out[r0 : r0 + 128, col_offset : col_offset + H * 128].float().reshape(128, H, 128).permute(1, 0, 2)
            got[b, :, t * 128 : (t + 1) * 128] = blk
    err = (got - ref).abs()
    tol = atol + ref.abs() / rdiv
    bad = err > tol
    assert not bad.any(), f"{int(bad.sum())}/{bad.numel()} mismatches, max err {err.max().item():.4g} at {torch.nonzero(bad)[0].tolist()}"
    if extra_cols or col_offset:
        other = torch.cat([out[:, :col_offset], out[:, col_offset + H * 128 :]], dim=1)
        assert torch.isnan(other.float()).all(), "attention wrote outside its column range"


def test_attention_single_tile():
    _run(1, 1, 128, 0, 0)


def test_attention_multi_tile_joint():
    _run(2, 2, 128, 256, 128)


def test_attention_large_logits_rescale():
    _run(1, 2, 128, 384, 256, qscale=6.0)


@pytest.mark.parametrize("mask_mode", [1, 2])
def test_attention_block_masks(mask_mode):
    _run(1, 2, 128, 256, 256, mask_mode=mask_mode)


def test_attention_c_factor_bias():
    _run(1, 2, 128, 128, 256, mask_mode=1, cross_bias=math.log(1.7))


def test_attention_strided_out():
    _run(1, 2, 128, 128, 128, col_offset=64, extra_cols=192)


@pytest.mark.parametrize("cfg", [
    # (B, H, nt, ni, nc, mask_mode, cross_bias, pretend-SM count)
    (2, 2, 128, 256, 128, 0, 0.0, 3), (2, 2, 128, 256, 128, 0, 0.0, 7), (1, 2, 128, 256, 256, 1, 0.0, 3),
    (1, 2, 128, 256, 256, 2, 0.0, 3), (1, 2, 512, 1024, 1024, 1, 0.0, 19), (1, 3, 128, 384, 256, 0, 0.0, 5),
    (1, 2, 128, 128, 256, 1, math.log(1.7), 3), (1, 3, 256, 256, 1536, 1, 0.0, 23), (1, 2, 128, 384, 256, 0, 0.0, 4)])
def test_attention_split_work_schedule(cfg):
    """The persistent schedule with more units than SMs: CTA ranges cut units at arbitrary KV iterations, the cut
    units are finished from fp32 partials exchanged through the workspace (1, 2 and 3 contributors per unit, one and
    two query tiles per unit, masks, bias).  Same tolerance as the unsplit kernel; run twice (flags return to idle)."""
    from loongx_b200 import _lib as L

    B, H, nt, ni, nc, mask_mode, cross_bias, ctas = cfg
    L.lib.lx_debug_attention_ctas(ctas)
    try:
        _run(B, H, nt, ni, nc, mask_mode=mask_mode, cross_bias=cross_bias)
        # peaked rows (|logit| ~ 4 sigma): a handful of bf16-rounded probabilities carry the row, so the output
        # inherits their 2^-9 relative rounding un-averaged: tolerance 8e-3 + 2^-6 |ref|
        _run(B, H, nt, ni, nc, mask_mode=mask_mode, cross_bias=cross_bias, qscale=4.0, atol=8e-3, rdiv=64)
    finally:
        L.lib.lx_debug_attention_ctas(0)


def test_attention_split_equals_unsplit_lse():
    """lse and the output rows of the split schedule against the unsplit one (no workspace use when the pretended SM
    count covers every unit): fp32 summation order differs, nothing else."""
    from loongx_b200 import _lib as L
    from loongx_b200 import ops

    B, H, nt, ni, nc = 1, 4, 256, 512, 256
    S = nt + ni + nc
    q, k, v = _mk((B, H, S, 128), 2.0, 1), _mk((B, H, S, 128), 1.0, 2), _mk((B, H, S, 128), 1.0, 3)
    orb = ops.make_out_row_base(B, nt, ni, nc, "cuda")
    res = []
    for ctas in (0, 5, 11):
        L.lib.lx_debug_attention_ctas(ctas)
        out = torch.zeros((B * S, H * 128), device="cuda", dtype=torch.bfloat16)
        lse = torch.zeros((B, H, S), device="cuda", dtype=torch.float32)
        ops.attention(q, k, v, out, orb, n_cond=nc, lse=lse)
        torch.cuda.synchronize()
        res.append((out.float(), lse))
    L.lib.lx_debug_attention_ctas(0)
    for out, lse in res[1:]:
        assert (out - res[0][0]).abs().max().item() <= 2e-2 * res[0][0].abs().max().item()
        assert _rel(out, res[0][0]) < 3e-3
        assert (lse - res[0][1]).abs().max().item() < 1e-3


def test_attention_flux_shape_throughput():
    """512x512 edit shape (S = 512 + 1024 + 1024, 24 heads): parity + printed TFLOP/s (informational)."""
    from loongx_b200 import ops

    B, H, nt, ni, nc = 1, 24, 512, 1024, 1024
    S = nt + ni + nc
    q, k, v = _mk((B, H, S, 128), 1.0, 1), _mk((B, H, S, 128), 1.0, 2), _mk((B, H, S, 128), 1.0, 3)
    out = torch.empty((B * S, H * 128), device="cuda", dtype=torch.bfloat16)
    orb = ops.make_out_row_base(B, nt, ni, nc, "cuda")
    for _ in range(3):
        ops.attention(q, k, v, out, orb, n_cond=nc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention(q, k, v, out, orb, n_cond=nc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"\n[attention smoke] B{B} H{H} S{S}: {ms:.3f} ms, {4 * B * H * S * S * 128 / ms / 1e9:.1f} TFLOP/s")
    ref = _ref(q, k, v, nc, 0, 0.0)  # explicit fp32 math (torch SDPA on sm_100 spends minutes in a JIT warm-up)
    got = out.reshape(S, H, 128).permute(1, 0, 2).float()  # B == 1: stream-major rows == sequence order
    err = (got - ref[0]).abs()
    assert (err <= 4e-3 + ref[0].abs() / 128).all(), err.max()


def _sdpa_ref_grads(q, k, v, d_out, n_cond, mask_mode, c_factor):
    import math

    import torch.nn.functional as F

    S = q.shape[2]
    mask = None
    if n_cond > 0:
        if c_factor is not None:
            mask = torch.zeros(S, S, device=q.device)
            mask[-n_cond:, :-n_cond] = math.log(c_factor)
            mask[:-n_cond, -n_cond:] = math.log(c_factor)
        elif mask_mode == 1:
            mask = torch.ones(S, S, device=q.device, dtype=torch.bool)
            mask[-n_cond:, :-n_cond] = False
            mask[:-n_cond, -n_cond:] = False
        elif mask_mode == 2:
            mask = torch.ones(S, S, device=q.device, dtype=torch.bool)
            mask[-n_cond:, :-n_cond] = False
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    o = F.scaled_dot_product_attention(qf, kf, vf, attn_mask=mask)
    return (o,) + torch.autograd.grad(o, (qf, kf, vf), d_out.float())


@pytest.mark.parametrize("B,H,nt,ni,nc,mask_mode,c_factor", [
    (1, 1, 128, 0, 0, 0, None), (2, 2, 128, 128, 128, 0, None), (1, 2, 128, 256, 256, 1, None),
    (1, 2, 128, 256, 256, 2, None), (1, 2, 256, 256, 128, 0, 1.7), (1, 3, 512, 1024, 1024, 0, None)])
def test_attention_backward_vs_autograd(B, H, nt, ni, nc, mask_mode, c_factor):
    """lx_attention (with lse) + lx_attention_bwd_prep + lx_attention_bwd against fp32 autograd of SDPA on the same
    bf16 inputs.  Tolerance: relL2 <= 2e-2 per gradient (bf16 P / dS operands, fp32 accumulation)."""
    import math

    from loongx_b200 import ops

    S = nt + ni + nc
    g = torch.Generator(device="cuda").manual_seed(B * 100 + S)
    mk = lambda *s, scale=1.0: (torch.randn(*s, generator=g, device="cuda") * scale).bfloat16()  # noqa: E731
    q, k, v, d_o = mk(B, H, S, 128), mk(B, H, S, 128), mk(B, H, S, 128), mk(B, H, S, 128)
    tm = ops.make_tile_meta(B, nt, ni, nc, "cuda")
    orb = ops.make_out_row_base(B, nt, ni, nc, "cuda")
    R, D = B * S, H * 128
    out_rows = torch.zeros(R, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, S, device="cuda")
    cb = math.log(c_factor) if c_factor is not None else 0.0
    ops.attention(q, k, v, out_rows, orb, n_cond=nc, mask_mode=mask_mode, cross_bias=cb, lse=lse)
    o_ref, dq_ref, dk_ref, dv_ref = _sdpa_ref_grads(q, k, v, d_o, nc, mask_mode, c_factor)
    # lse (log2 domain) against the reference softmax normaliser
    sc = 1.0 / math.sqrt(128)
    logits = torch.einsum("bhqd,bhkd->bhqk", q.float(), k.float()) * sc
    if nc > 0 and c_factor is not None:
        logits[..., -nc:, :-nc] += cb
        logits[..., :-nc, -nc:] += cb
    elif nc > 0 and mask_mode == 1:
        logits[..., -nc:, :-nc] = -float("inf")
        logits[..., :-nc, -nc:] = -float("inf")
    elif nc > 0 and mask_mode == 2:
        logits[..., -nc:, :-nc] = -float("inf")
    lse_ref = torch.logsumexp(logits, -1) / math.log(2.0)
    assert (lse - lse_ref).abs().max().item() < 2e-2
    # dO in the stream-major row layout (what the training step holds) -> prep -> backward
    def to_rows(x):  # [B,H,S,128] -> [R, D]
        xs = x.permute(0, 2, 1, 3).reshape(B, S, D)
        return torch.cat([xs[:, :nt].reshape(B * nt, D), xs[:, nt:nt + ni].reshape(B * ni, D), xs[:, nt + ni:].reshape(B * nc, D)])
    d_rows = to_rows(d_o).contiguous()
    d_heads = torch.zeros_like(d_o)
    delta = torch.zeros(B, H, S, device="cuda")
    ops.attention_bwd_prep(d_rows, out_rows, H, tm, d_heads, delta)
    assert torch.equal(d_heads, d_o)
    delta_ref = (d_o.float() * o_ref).sum(-1)
    assert _rel(delta, delta_ref) < 2e-2
    dq = torch.zeros(B, H, S, 128, device="cuda")
    dk, dv = torch.zeros_like(q), torch.zeros_like(q)
    ops.attention_bwd(q, k, v, d_heads, lse, delta, dq, dk, dv, n_cond=nc, mask_mode=mask_mode, cross_bias=cb)
    torch.cuda.synchronize()
    e = (_rel(dq, dq_ref), _rel(dk, dk_ref), _rel(dv, dv_ref))
    print(f"\n[attn bwd B{B} H{H} S{S} mask{mask_mode} cf{c_factor}] relL2 dq {e[0]:.4g} dk {e[1]:.4g} dv {e[2]:.4g}")
    assert max(e) < 2e-2, e


def test_attention_backward_flux_shape_throughput():
    from loongx_b200 import ops

    B, H, S = 1, 24, 2560
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda: torch.randn(B, H, S, 128, generator=g, device="cuda").bfloat16()  # noqa: E731
    q, k, v, d_o = mk(), mk(), mk(), mk()
    lse = torch.randn(B, H, S, device="cuda") + 12.0
    delta = torch.randn(B, H, S, device="cuda")
    dq = torch.zeros(B, H, S, 128, device="cuda")
    dk, dv = torch.zeros_like(q), torch.zeros_like(q)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.attention_bwd(q, k, v, d_o, lse, delta, dq, dk, dv, n_cond=1024)
    e0.record()
    for _ in range(10):
        ops.attention_bwd(q, k, v, d_o, lse, delta, dq, dk, dv, n_cond=1024)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"\n[attn bwd] S=2560 H=24: {ms:.3f} ms  {10.0 * B * H * S * S * 128 / ms / 1e9:.0f} TFLOP/s")
    assert torch.isfinite(dk.float()).all()


@pytest.mark.parametrize("nt,ni,nc,mask_mode", [(100, 200, 130, 0), (77, 128, 0, 0), (128, 250, 250, 2), (1, 129, 127, 1)])
def test_attention_ragged_streams_forward_and_backward(nt, ni, nc, mask_mode):
    """Ragged stream lengths: every stream padded to a multiple of 128 tokens, padding keys masked inside the kernels
    (junk values in the padding must not leak).  Forward output rows and dq/dk/dv of the valid tokens vs fp32 autograd
    on the un-padded tensors."""
    from loongx_b200 import ops

    B, H = 2, 2
    pad = lambda n: (n + 127) // 128 * 128  # noqa: E731
    ntp, nip, ncp = pad(nt), pad(ni), pad(nc)
    Sp, S = ntp + nip + ncp, nt + ni + nc
    g = torch.Generator(device="cuda").manual_seed(nt * 7 + ni)
    mk = lambda *s: torch.randn(*s, generator=g, device="cuda").bfloat16()  # noqa: E731
    qp, kp, vp, dop = mk(B, H, Sp, 128), mk(B, H, Sp, 128), mk(B, H, Sp, 128) * 3, mk(B, H, Sp, 128)
    valid = torch.cat([torch.arange(0, nt), torch.arange(ntp, ntp + ni), torch.arange(ntp + nip, ntp + nip + nc)]).cuda()
    q, k, v, d_o = (t[:, :, valid].contiguous() for t in (qp, kp, vp, dop))
    o_ref, dq_ref, dk_ref, dv_ref = _sdpa_ref_grads(q, k, v, d_o, nc, mask_mode, None)
    pads = (ntp - nt, nip - ni, ncp - nc)
    tm = ops.make_tile_meta(B, ntp, nip, ncp, "cuda")
    orb = ops.make_out_row_base(B, ntp, nip, ncp, "cuda")
    D = H * 128
    out_rows = torch.zeros(B * Sp, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B, H, Sp, device="cuda")
    ops.attention(qp, kp, vp, out_rows, orb, n_cond=ncp, mask_mode=mask_mode, lse=lse, pads=pads, n_txt=ntp)

    def rows_to_seq(rows):  # stream-major padded rows [B*Sp, D] -> [B, H, Sp, 128]
        rt, ri = B * ntp, B * nip
        xs = torch.cat([rows[:rt].view(B, ntp, D), rows[rt:rt + ri].view(B, nip, D), rows[rt + ri:].view(B, ncp, D)], 1)
        return xs.view(B, Sp, H, 128).permute(0, 2, 1, 3)

    def seq_to_rows(x):  # [B,H,Sp,128] -> [B*Sp, D]
        xs = x.permute(0, 2, 1, 3).reshape(B, Sp, D)
        return torch.cat([xs[:, :ntp].reshape(B * ntp, D), xs[:, ntp:ntp + nip].reshape(B * nip, D), xs[:, ntp + nip:].reshape(B * ncp, D)])

    got = rows_to_seq(out_rows)[:, :, valid]
    e_o = _rel(got, o_ref)
    # backward: dO of the padding query rows is zero (what the training step guarantees)
    dop_z = torch.zeros_like(dop)
    dop_z[:, :, valid] = d_o
    d_heads = torch.zeros_like(dop)
    delta = torch.zeros(B, H, Sp, device="cuda")
    ops.attention_bwd_prep(seq_to_rows(dop_z).contiguous(), out_rows, H, tm, d_heads, delta)
    dq = torch.zeros(B, H, Sp, 128, device="cuda")
    dk, dv = torch.zeros_like(qp), torch.zeros_like(qp)
    ops.attention_bwd(qp, kp, vp, d_heads, lse, delta, dq, dk, dv, n_cond=ncp, mask_mode=mask_mode, pads=pads, n_txt=ntp)
    torch.cuda.synchronize()
    e = (_rel(dq[:, :, valid], dq_ref), _rel(dk[:, :, valid], dk_ref), _rel(dv[:, :, valid], dv_ref))
    inv = torch.ones(Sp, dtype=torch.bool, device="cuda")
    inv[valid] = False
    print(f"\n[ragged attn {nt}/{ni}/{nc} mask{mask_mode}] out relL2 {e_o:.4g}; dq {e[0]:.4g} dk {e[1]:.4g} dv {e[2]:.4g}")
    assert e_o < 1e-2 and max(e) < 2e-2
    if inv.any():  # padding keys receive no gradient
        assert float(dk[:, :, inv].abs().max()) == 0.0 and float(dv[:, :, inv].abs().max()) == 0.0
