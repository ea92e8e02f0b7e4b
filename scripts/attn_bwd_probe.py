"""Attention-backward timing probe (development aid): S=2560, H=24, B in argv; optional debug flags."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loongx_b200 import ops, _lib as L

H, S = 24, 2560
for B in [int(a) for a in sys.argv[1:]] or [1, 4]:
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda: torch.randn(B, H, S, 128, generator=g, device="cuda").bfloat16()
    q, k, v, d_o = mk(), mk(), mk(), mk()
    lse = torch.randn(B, H, S, device="cuda") + 12.0
    delta = torch.randn(B, H, S, device="cuda")
    dq = torch.zeros(B, H, S, 128, device="cuda")
    dk, dv = torch.zeros_like(q), torch.zeros_like(q)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for flags in (0, 1):
        L.lib.lx_attention_bwd_debug_flags(flags)
        for _ in range(3):
            ops.attention_bwd(q, k, v, d_o, lse, delta, dq, dk, dv, n_cond=1024)
        e0.record()
        for _ in range(10):
            ops.attention_bwd(q, k, v, d_o, lse, delta, dq, dk, dv, n_cond=1024)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"B={B} flags={flags}: {ms:.3f} ms  {10.0 * B * H * S * S * 128 / ms / 1e9:.0f} TFLOP/s", flush=True)
    # phase clocks of the first compute thread of every CTA (flags bit 1)
    import ctypes
    prof = torch.zeros(B * H * (S // 128), 4, dtype=torch.int64, device="cuda")
    L.lib.lx_attention_bwd_debug_prof.argtypes = [ctypes.c_void_p]
    L.lib.lx_attention_bwd_debug_prof(prof.data_ptr())
    L.lib.lx_attention_bwd_debug_flags(2)
    ops.attention_bwd(q, k, v, d_o, lse, delta, dq, dk, dv, n_cond=1024)
    torch.cuda.synchronize()
    L.lib.lx_attention_bwd_debug_flags(0)
    L.lib.lx_attention_bwd_debug_prof(None)
    m = prof.double().mean(0) / (S // 128)
    print(f"B={B} clocks per (key tile, query tile) iteration, CTA mean: wait S^T/dP^T {m[0]:.0f}, softmax phase {m[1]:.0f}, "
          f"wait dQ {m[2]:.0f}, dQ read-back {m[3]:.0f}, total {m.sum():.0f}", flush=True)
