"""lx_gemm_bf16 against torch.matmul (cuBLAS) on the DiT's GEMM shapes, both back to back with rotating weight copies
(development aid: is the vendor library any faster on these M = 2560 shapes?)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from loongx_b200 import ops

shapes = [("qkv", 9216, 3072), ("attn_out", 3072, 3072), ("ff_up", 12288, 3072), ("ff_down", 3072, 12288),
          ("single_qkv_mlp", 21504, 3072), ("single_out", 3072, 15360)]
M = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, N, K in shapes:
    ncopy = max(2, (300 << 20) // (N * K * 2) + 1)
    Ws = [torch.randn(N, K, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(ncopy)]
    A = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = []
    for which in ("lx", "cublas"):
        f = (lambda i: ops.gemm(A, Ws[i % ncopy], None, out)) if which == "lx" else (lambda i: torch.matmul(A, Ws[i % ncopy].t(), out=out))
        for i in range(ncopy):
            f(i)
        reps = 3 * ncopy
        e0.record()
        for i in range(reps):
            f(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res.append(f"{which} {ms * 1e3:7.1f} us {2.0 * M * N * K / ms / 1e9:7.1f} TF")
    print(f"M={M} {name:15s} N={N:6d} K={K:6d} | " + "   ".join(res), flush=True)
    del Ws
