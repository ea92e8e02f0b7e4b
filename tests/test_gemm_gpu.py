"""GPU parity tests of the tcgen05 GEMM and its fused epilogues against plain torch fp32 math.

Tolerance (stated): inputs are bf16, accumulation fp32; the output is rounded once to bf16, so
|err| <= 2^-8 * |ref| + small absolute slack from accumulation-order differences.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, generator=g, device="cuda", dtype=torch.float32) * scale).to(dtype)


def _close(out, ref, rtol=1.0 / 128, atol=2e-2, what=""):
    out = out.float()
    err = (out - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    assert not bad.any(), (
        f"{what}: {int(bad.sum())} / {bad.numel()} mismatches, max err {err.max().item():.4g}, "
        f"first at {tuple(torch.nonzero(bad)[0].tolist())}"
    )


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 256), (384, 768, 3136), (200, 328, 200), (2560, 3072, 3072)])
def test_gemm_bias(M, N, K):
    from loongx_b200 import ops, _lib as L

    A, W = _mk((M, K), 1.0, 1), _mk((N, K), 0.05, 2)
    bias = _mk((N,), 1.0, 3, torch.float32)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, bias, out, L.EPI_BIAS)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias
    _close(out, ref, what=f"bias {M}x{N}x{K}")


@pytest.mark.parametrize("tile_n", [256, 224, 192, 128])
def test_gemm_tile_variants(tile_n):
    """Every N-tile variant of the kernel (wave-quantisation tuning) gives the same answer, incl. ragged M/N/K."""
    from loongx_b200 import ops, _lib as L

    for (M, N, K) in [(384, 3072, 320), (200, 456, 136), (640, 128, 1152), (100, 8, 192)]:
        A, W = _mk((M, K), 1.0, 41), _mk((N, K), 0.05, 42)
        bias = _mk((N,), 1.0, 43, torch.float32)
        out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.gemm(A, W, bias, out, L.EPI_BIAS_GELU, tile_n=tile_n)
        torch.cuda.synchronize()
        ref = torch.nn.functional.gelu(A.float() @ W.float().t() + bias, approximate="tanh")
        _close(out, ref, what=f"tile_n={tile_n} {M}x{N}x{K}")


def test_gemm_row_groups():
    """Three row groups (text / image / condition rows) with their own weight panel and bias in one launch."""
    from loongx_b200 import ops, _lib as L

    M, N, K = 640, 512, 256
    A = _mk((M, K), 1.0, 50)
    Ws = [_mk((N, K), 0.05, 51 + i) for i in range(3)]
    bs = [_mk((N,), 1.0, 54 + i, torch.float32) for i in range(3)]
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, Ws[0], bs[0], out, L.EPI_BIAS, groups=[(Ws[1], bs[1], 128), (Ws[2], bs[2], 384)])
    torch.cuda.synchronize()
    ref = torch.cat([A[:128].float() @ Ws[0].float().t() + bs[0], A[128:384].float() @ Ws[1].float().t() + bs[1],
                     A[384:].float() @ Ws[2].float().t() + bs[2]])
    _close(out, ref, what="row groups")


def test_gemm_strided_ext_and_f32():
    """A / W with row stride > K (the LoRA K-extension layout) and an fp32 output segment."""
    from loongx_b200 import ops, _lib as L

    M, N, K, ld = 256, 256, 128 + 64, 320
    Abuf, Wbuf = _mk((M, ld), 1.0, 4), _mk((N, ld), 0.1, 5)
    out = torch.zeros((M, N), device="cuda", dtype=torch.float32)
    ops.gemm(Abuf, Wbuf, None, out, L.EPI_BIAS_F32, K=K)
    torch.cuda.synchronize()
    ref = Abuf[:, :K].float() @ Wbuf[:, :K].float().t()
    assert torch.allclose(out, ref, rtol=1e-3, atol=1e-2), (out - ref).abs().max()


@pytest.mark.parametrize("mode", ["gelu", "silu"])
def test_gemm_activation(mode):
    from loongx_b200 import ops, _lib as L

    M, N, K = 256, 512, 512
    A, W = _mk((M, K), 1.0, 6), _mk((N, K), 0.06, 7)
    bias = _mk((N,), 0.5, 8, torch.float32)
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, bias, out, L.EPI_BIAS_GELU if mode == "gelu" else L.EPI_BIAS_SILU)
    torch.cuda.synchronize()
    pre = A.float() @ W.float().t() + bias
    ref = torch.nn.functional.gelu(pre, approximate="tanh") if mode == "gelu" else torch.nn.functional.silu(pre)
    _close(out, ref, what=mode)


def test_gemm_gate_residual_inplace():
    from loongx_b200 import ops, _lib as L

    B, nt, ni, nc, D, K = 2, 128, 256, 128, 512, 256
    S = nt + ni + nc
    R = B * S
    meta = ops.make_tile_meta(B, nt, ni, nc, "cuda")
    A, W = _mk((R, K), 1.0, 9), _mk((D, K), 0.06, 10)
    bias = _mk((D,), 0.5, 11, torch.float32)
    x = _mk((R, D), 1.0, 12)
    gates = [_mk((B, 3 * D), 1.0, 13 + i) for i in range(3)]  # gate = columns [D, 2D) of a modulation row
    gviews = [g[:, D : 2 * D] for g in gates]
    x0 = x.clone()
    ops.gemm(A, W, bias, x, L.EPI_GATE_RESIDUAL, tile_meta=meta, residual=x, gate=gviews)
    torch.cuda.synchronize()
    lin = A.float() @ W.float().t() + bias
    ref = torch.empty_like(lin)
    m = meta.cpu().tolist()
    for t, (stream, b, _, _) in enumerate(m):
        sl = slice(t * 128, (t + 1) * 128)
        ref[sl] = x0[sl].float() + gviews[stream][b].float()[None, :] * lin[sl]
    _close(x, ref, what="gate_residual")


def _rope_table(S, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ang = torch.rand((S, 64), generator=g, device="cuda") * 6.28
    return torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()  # [S, 64, 2]


def _ref_qkv(lin, meta, B, H, S, rms_q, rms_k, rope, eps):
    """lin [R, 3D] fp32 -> q,k,v [B,H,S,128] following block.py:34-41,74-78 (+ diffusers RMSNorm / apply_rotary_emb)."""
    D = H * 128
    outs = [torch.zeros((B, H, S, 128), device=lin.device) for _ in range(3)]
    for t, (stream, b, seq_row, _) in enumerate(meta.cpu().tolist()):
        s0 = seq_row % S
        rows = lin[t * 128 : (t + 1) * 128]
        for sec in range(3):
            x = rows[:, sec * D : (sec + 1) * D].reshape(128, H, 128).permute(1, 0, 2)  # [H,128,128]
            if sec < 2:
                w = (rms_q if sec == 0 else rms_k)[stream]
                x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w
                cos, sin = rope[s0 : s0 + 128, :, 0], rope[s0 : s0 + 128, :, 1]
                xr, xi = x[..., 0::2], x[..., 1::2]
                x = torch.stack([xr * cos - xi * sin, xi * cos + xr * sin], dim=-1).flatten(-2)
            outs[sec][b, :, s0 : s0 + 128] = x
    return outs


def test_gemm_qkv_epilogue_and_split_gelu():
    """Fused single-block projection: columns [0,3D) -> QKV epilogue, [3D,3D+F) -> GELU into a wider buffer."""
    from loongx_b200 import ops, _lib as L

    B, nt, ni, nc, H, K, F = 2, 128, 128, 256, 2, 320, 512
    D = H * 128
    S = nt + ni + nc
    R = B * S
    meta = ops.make_tile_meta(B, nt, ni, nc, "cuda")
    A, W = _mk((R, K), 1.0, 20), _mk((3 * D + F, K), 0.06, 21)
    bias = _mk((3 * D + F,), 0.3, 22, torch.float32)
    rms_q = [(_mk((128,), 0.2, 23 + i, torch.float32) + 1.0) for i in range(3)]
    rms_k = [(_mk((128,), 0.2, 26 + i, torch.float32) + 1.0) for i in range(3)]
    rope = _rope_table(S, 29)
    q = torch.full((B, H, S, 128), float("nan"), device="cuda", dtype=torch.bfloat16)
    k, v = q.clone(), q.clone()
    cat = torch.zeros((R, D + F + 64), device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, bias, None, L.EPI_QKV, n_split=3 * D, seg1=(L.EPI_BIAS_GELU, cat, D), tile_meta=meta,
             qkv=(q, k, v), rms_q=rms_q, rms_k=rms_k, rope=rope, rms_eps=1e-6)
    torch.cuda.synchronize()
    lin = A.float() @ W.float().t() + bias
    rq, rk, rv = _ref_qkv(lin[:, : 3 * D], meta, B, H, S, rms_q, rms_k, rope, 1e-6)
    _close(q, rq, what="q")
    _close(k, rk, what="k")
    _close(v, rv, what="v")
    ref_mlp = torch.nn.functional.gelu(lin[:, 3 * D :], approximate="tanh")
    _close(cat[:, D : D + F], ref_mlp, what="mlp segment")
    assert (cat[:, :D] == 0).all() and (cat[:, D + F :] == 0).all(), "GELU segment wrote outside its columns"


def test_gemm_training_epilogues():
    """The epilogues the training step adds (model.py:569-729 differentiated by hand): the single block's fused projection
    with the pre-norm q|k|v kept (qkv_pre) and the MLP segment written twice (gelu' for the backward, GELU for the
    forward: BIAS_GELU_DUAL); dX through the GELU as one multiply (MUL_AUX); gate + residual that also keeps the pre-gate
    projection (GATE_RESIDUAL with out2)."""
    from loongx_b200 import ops, _lib as L

    B, nt, ni, nc, H, K, F = 2, 128, 128, 256, 2, 320, 512
    D = H * 128
    S = nt + ni + nc
    R = B * S
    meta = ops.make_tile_meta(B, nt, ni, nc, "cuda")
    A, W = _mk((R, K), 1.0, 20), _mk((3 * D + F, K), 0.06, 21)
    bias = _mk((3 * D + F,), 0.3, 22, torch.float32)
    rms_q = [(_mk((128,), 0.2, 23 + i, torch.float32) + 1.0) for i in range(3)]
    rms_k = [(_mk((128,), 0.2, 26 + i, torch.float32) + 1.0) for i in range(3)]
    rope = _rope_table(S, 29)
    q = torch.full((B, H, S, 128), float("nan"), device="cuda", dtype=torch.bfloat16)
    k, v = q.clone(), q.clone()
    qm = torch.zeros((R, 3 * D + F + 64), device="cuda", dtype=torch.bfloat16)  # [q | k | v pre-norm | gelu' | guard]
    cat = torch.zeros((R, D + F + 64), device="cuda", dtype=torch.bfloat16)
    ops.gemm(A, W, bias, None, L.EPI_QKV, n_split=3 * D, seg1=(L.EPI_BIAS_GELU_DUAL, qm, 3 * D), out2=(cat, D),
             tile_meta=meta, qkv=(q, k, v), rms_q=rms_q, rms_k=rms_k, rope=rope, rms_eps=1e-6, qkv_pre=qm)
    torch.cuda.synchronize()
    lin = A.float() @ W.float().t() + bias
    rq, rk, rv = _ref_qkv(lin[:, : 3 * D], meta, B, H, S, rms_q, rms_k, rope, 1e-6)
    _close(q, rq, what="q")
    _close(k, rk, what="k")
    _close(v, rv, what="v")
    _close(qm[:, : 3 * D], lin[:, : 3 * D], what="qkv_pre")
    pre = lin[:, 3 * D :].clone().requires_grad_(True)
    act = torch.nn.functional.gelu(pre, approximate="tanh")
    (dact,) = torch.autograd.grad(act.sum(), pre)
    _close(cat[:, D : D + F], act, what="GELU of the dual segment")
    _close(qm[:, 3 * D : 3 * D + F], dact, what="gelu' of the dual segment")
    assert (cat[:, :D] == 0).all() and (cat[:, D + F :] == 0).all() and (qm[:, 3 * D + F :] == 0).all()
    # MUL_AUX: [plain | times the stored factor] like the single block's dCat (proj_out^T)
    G, Wt = _mk((R, K), 1.0, 31), _mk((256 + F, K), 0.06, 32)
    d_cat = torch.zeros((R, 256), device="cuda", dtype=torch.bfloat16)
    d_big = torch.zeros((R, 3 * D + F), device="cuda", dtype=torch.bfloat16)
    ops.gemm(G, Wt, None, d_cat, L.EPI_BIAS, n_split=256, seg1=(L.EPI_MUL_AUX, d_big, 3 * D), residual=qm)
    torch.cuda.synchronize()
    lin2 = G.float() @ Wt.float().t()
    _close(d_cat, lin2[:, :256], what="plain segment")
    _close(d_big[:, 3 * D :], lin2[:, 256:] * qm[:, 3 * D : 3 * D + F].float(), what="MUL_AUX segment")
    assert (d_big[:, : 3 * D] == 0).all()
    # GATE_RESIDUAL + out2
    A3, W3 = _mk((R, K), 1.0, 9), _mk((D, K), 0.06, 10)
    b3 = _mk((D,), 0.5, 11, torch.float32)
    x = _mk((R, D), 1.0, 12)
    gates = [_mk((B, 3 * D), 1.0, 13 + i)[:, D : 2 * D] for i in range(3)]
    x0, y = x.clone(), torch.zeros((R, D + 8), device="cuda", dtype=torch.bfloat16)
    ops.gemm(A3, W3, b3, x, L.EPI_GATE_RESIDUAL, tile_meta=meta, residual=x, gate=gates, out2=(y, 8))
    torch.cuda.synchronize()
    lin3 = A3.float() @ W3.float().t() + b3
    ref = torch.empty_like(lin3)
    for t, (stream, b, _, _) in enumerate(meta.cpu().tolist()):
        sl = slice(t * 128, (t + 1) * 128)
        ref[sl] = x0[sl].float() + gates[stream][b].float()[None, :] * lin3[sl]
    _close(x, ref, what="gate_residual")
    _close(y[:, 8:], lin3, what="pre-gate projection (out2)")
    assert (y[:, :8] == 0).all()


@pytest.mark.parametrize("sms", [6, 8, 12, 20])
def test_gemm_stream_k_head(sms):
    """The stream-K head (partial last wave cut along K, fp32 partial tiles exchanged through the workspace) on a
    pretended SM count so that small shapes take it: plain / ragged, three row groups, gate-residual in place, and the
    QKV + GELU two-segment epilogue; every case twice (the exchange flags must return to idle)."""
    import ctypes

    from loongx_b200 import ops, _lib as L

    L.lib.lx_debug_gemm_stream_k_launches.restype = ctypes.c_longlong
    L.lib.lx_debug_gemm_sms(sms)
    L.lib.lx_debug_gemm_stream_k(1)  # off by default (slower than the partial wave it removes at the DiT's shapes)
    n0 = L.lib.lx_debug_gemm_stream_k_launches()
    try:
        for rep in range(2):
            # plain + ragged edges
            for (M, N, K) in [(640, 1280, 1024), (600, 1096, 1032), (1280, 768, 2048)]:
                A, W = _mk((M, K), 1.0, 61), _mk((N, K), 0.05, 62)
                bias = _mk((N,), 1.0, 63, torch.float32)
                out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
                ops.gemm(A, W, bias, out, L.EPI_BIAS_GELU)
                ref = torch.nn.functional.gelu(A.float() @ W.float().t() + bias, approximate="tanh")
                _close(out, ref, atol=3e-2, what=f"stream-k gelu {M}x{N}x{K} sms={sms}")
            # three row groups
            M, N, K = 1280, 768, 1536
            A = _mk((M, K), 1.0, 50)
            Ws = [_mk((N, K), 0.05, 51 + i) for i in range(3)]
            bs = [_mk((N,), 1.0, 54 + i, torch.float32) for i in range(3)]
            out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
            ops.gemm(A, Ws[0], bs[0], out, L.EPI_BIAS, groups=[(Ws[1], bs[1], 256), (Ws[2], bs[2], 768)])
            ref = torch.cat([A[:256].float() @ Ws[0].float().t() + bs[0], A[256:768].float() @ Ws[1].float().t() + bs[1],
                             A[768:].float() @ Ws[2].float().t() + bs[2]])
            _close(out, ref, atol=3e-2, what=f"stream-k row groups sms={sms}")
            # gate * y + residual, in place
            B, nt, ni, nc, D, K = 2, 256, 512, 256, 1280, 2048
            R = B * (nt + ni + nc)
            meta = ops.make_tile_meta(B, nt, ni, nc, "cuda")
            A, W = _mk((R, K), 1.0, 9), _mk((D, K), 0.03, 10)
            bias = _mk((D,), 0.5, 11, torch.float32)
            x = _mk((R, D), 1.0, 12)
            gates = [_mk((B, 3 * D), 1.0, 13 + i) for i in range(3)]
            gviews = [g[:, D: 2 * D] for g in gates]
            x0 = x.clone()
            ops.gemm(A, W, bias, x, L.EPI_GATE_RESIDUAL, tile_meta=meta, residual=x, gate=gviews)
            lin = A.float() @ W.float().t() + bias
            ref = torch.empty_like(lin)
            for t, (stream, b, _, _) in enumerate(meta.cpu().tolist()):
                sl = slice(t * 128, (t + 1) * 128)
                ref[sl] = x0[sl].float() + gviews[stream][b].float()[None, :] * lin[sl]
            _close(x, ref, atol=3e-2, what=f"stream-k gate_residual sms={sms}")
            # QKV epilogue + GELU segment
            B, nt, ni, nc, H, K, F = 2, 256, 256, 256, 2, 1024, 512
            D = H * 128
            S = nt + ni + nc
            R = B * S
            meta = ops.make_tile_meta(B, nt, ni, nc, "cuda")
            A, W = _mk((R, K), 1.0, 20), _mk((3 * D + F, K), 0.04, 21)
            bias = _mk((3 * D + F,), 0.3, 22, torch.float32)
            rms_q = [(_mk((128,), 0.2, 23 + i, torch.float32) + 1.0) for i in range(3)]
            rms_k = [(_mk((128,), 0.2, 26 + i, torch.float32) + 1.0) for i in range(3)]
            rope = _rope_table(S, 29)
            q = torch.full((B, H, S, 128), float("nan"), device="cuda", dtype=torch.bfloat16)
            k, v = q.clone(), q.clone()
            cat = torch.zeros((R, D + F + 64), device="cuda", dtype=torch.bfloat16)
            ops.gemm(A, W, bias, None, L.EPI_QKV, n_split=3 * D, seg1=(L.EPI_BIAS_GELU, cat, D), tile_meta=meta,
                     qkv=(q, k, v), rms_q=rms_q, rms_k=rms_k, rope=rope, rms_eps=1e-6)
            lin = A.float() @ W.float().t() + bias
            rq, rk, rv = _ref_qkv(lin[:, : 3 * D], meta, B, H, S, rms_q, rms_k, rope, 1e-6)
            _close(q, rq, what=f"stream-k q sms={sms}")
            _close(k, rk, what=f"stream-k k sms={sms}")
            _close(v, rv, atol=3e-2, what=f"stream-k v sms={sms}")
            _close(cat[:, D: D + F], torch.nn.functional.gelu(lin[:, 3 * D:], approximate="tanh"), atol=3e-2,
                   what=f"stream-k mlp segment sms={sms}")
        torch.cuda.synchronize()
        assert L.lib.lx_debug_gemm_stream_k_launches() > n0, "no launch took the stream-K path: the test shapes are stale"
    finally:
        L.lib.lx_debug_gemm_sms(0)
        L.lib.lx_debug_gemm_stream_k(0)


def test_gemm_stream_k_on_off_agree_at_flux_shapes():
    """The DiT's N = 3072 projections at M = 2560 with and without the stream-K head (same kernel, different schedule):
    results agree to fp32 summation order; prints both timings (informational)."""
    from loongx_b200 import ops, _lib as L

    for (N, K) in [(3072, 3072), (3072, 12288), (12288, 3072)]:
        M = 2560
        A, W = _mk((M, K), 1.0, 71), _mk((N, K), 0.02, 72)
        outs, times = [], []
        for on in (0, 1):
            L.lib.lx_debug_gemm_stream_k(on)
            out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
            for _ in range(3):
                ops.gemm(A, W, None, out, L.EPI_BIAS)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.gemm(A, W, None, out, L.EPI_BIAS)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1) / 20)
            outs.append(out.float())
        L.lib.lx_debug_gemm_stream_k(0)
        tf = [2 * M * N * K / t / 1e9 for t in times]
        print(f"\n[gemm stream-k] {M}x{N}x{K}: off {times[0]*1e3:.1f} us ({tf[0]:.0f} TF), on {times[1]*1e3:.1f} us ({tf[1]:.0f} TF)")
        assert ((outs[0] - outs[1]).abs() <= 2e-2 + outs[0].abs() / 128).all()


def test_gemm_bad_args_raise():
    from loongx_b200 import ops, _lib as L

    A, W = _mk((128, 60), 1.0, 1), _mk((256, 60), 1.0, 2)  # K not a multiple of 8
    out = torch.empty((128, 256), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(L.LoongXNativeError):
        ops.gemm(A, W, None, out, L.EPI_BIAS)


def test_gemm_throughput_smoke():
    """Not a benchmark: prints achieved TFLOP/s of the plain GEMM so the first GPU run shows where we stand."""
    from loongx_b200 import ops, _lib as L

    M, N, K = 2560, 12288, 3072
    A, W = _mk((M, K), 1.0, 30), _mk((N, K), 0.02, 31)
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        ops.gemm(A, W, None, out, L.EPI_BIAS)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(A, W, None, out, L.EPI_BIAS)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tf = 2 * M * N * K / ms / 1e9
    print(f"\n[gemm smoke] {M}x{N}x{K}: {ms:.3f} ms, {tf:.1f} TFLOP/s")
    ref = A.float() @ W.float().t()
    _close(out, ref, what="throughput shape")
    assert math.isfinite(tf)
