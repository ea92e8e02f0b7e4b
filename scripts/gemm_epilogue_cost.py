"""How much do the fused epilogues cost?  Same shape, plain bias vs gate-residual (out / ff_down / proj_out) and vs the
QKV epilogue (RMSNorm + RoPE + head scatter).  Weight panels rotate through > 256 MB so W always streams from HBM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loongx_b200 import ops, _lib as L

dev = "cuda"
B, nt, ni, nc = 1, 512, 1024, 1024
M = B * (nt + ni + nc)
tm = ops.make_tile_meta(B, nt, ni, nc, dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def bench(fn, ncopy):
    for i in range(ncopy):
        fn(i)
    reps = 4 * ncopy
    e0.record()
    for i in range(reps):
        fn(i % ncopy)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, N, K in [("out", 3072, 3072), ("ff_down", 3072, 12288), ("proj_out", 3072, 15360), ("qkv", 9216, 3072)]:
    ncopy = max(2, (300 << 20) // (N * K * 2) + 1)
    Ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(ncopy)]
    bias = torch.zeros(N, device=dev)
    A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    flop = 2.0 * M * N * K
    t_bias = bench(lambda i: ops.gemm(A, Ws[i], bias, out, L.EPI_BIAS), ncopy)
    line = f"{name:9s} N={N:5d} K={K:5d}  bias {t_bias*1e3:7.1f} us {flop/t_bias/1e9:6.0f} TF"
    if N == 3072:
        X = torch.randn(M, N, device=dev, dtype=torch.bfloat16)
        gate = [torch.randn(B, N, device=dev, dtype=torch.bfloat16) for _ in range(3)]
        t_g = bench(lambda i: ops.gemm(A, Ws[i], bias, X, L.EPI_GATE_RESIDUAL, tile_meta=tm, residual=X, gate=gate), ncopy)
        line += f" | gate+residual {t_g*1e3:7.1f} us {flop/t_g/1e9:6.0f} TF ({(t_g/t_bias-1)*100:+.1f} %)"
    else:
        H, S = 24, nt + ni + nc
        q, k, v = (torch.empty(B, H, S, 128, device=dev, dtype=torch.bfloat16) for _ in range(3))
        rope = torch.rand(S, 64, 2, device=dev)
        w = torch.ones(128, device=dev)
        t_q = bench(lambda i: ops.gemm(A, Ws[i], bias, None, L.EPI_QKV, tile_meta=tm, qkv=(q, k, v), rms_q=[w, w, w],
                                       rms_k=[w, w, w], rope=rope), ncopy)
        line += f" | qkv epilogue {t_q*1e3:7.1f} us {flop/t_q/1e9:6.0f} TF ({(t_q/t_bias-1)*100:+.1f} %)"
    print(line, flush=True)
    del Ws
