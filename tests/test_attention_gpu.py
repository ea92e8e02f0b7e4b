"""GPU parity tests of the tcgen05 joint-attention kernel against fp32 torch math (block.py:106-135 semantics).

Tolerance (stated): P is rounded to bf16 before the second GEMM and the output is rounded to bf16:
|err| <= 4e-3 + 2^-7 |ref|.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


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


def _run(B, H, nt, ni, nc, mask_mode=0, cross_bias=0.0, qscale=1.0, col_offset=0, extra_cols=0):
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
    tol = 4e-3 + ref.abs() / 128
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
