"""Does the N-tile picker (pick_tile_n) choose the fastest variant?  Per DiT GEMM shape: auto vs forced 256 / 224 / 192,
plain-bias epilogue, weights rotating through > 256 MB.  (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loongx_b200 import ops

dev = "cuda"
shapes = [("qkv", 9216, 3072), ("attn_out", 3072, 3072), ("ff_up", 12288, 3072), ("ff_down", 3072, 12288),
          ("single_qkv_mlp", 21504, 3072), ("single_out", 3072, 15360)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for M in [int(a) for a in sys.argv[1:]] or [2560, 10240]:
    for name, N, K in shapes:
        ncopy = max(2, (300 << 20) // (N * K * 2) + 1)
        Ws = [torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.02 for _ in range(ncopy)]
        A = torch.randn(M, K, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        res = []
        for rep in range(2):
            for tn in (0, 256, 224, 192):
                reps = 4 * ncopy
                for i in range(ncopy):
                    ops.gemm(A, Ws[i], None, out, tile_n=tn)
                e0.record()
                for i in range(reps):
                    ops.gemm(A, Ws[i % ncopy], None, out, tile_n=tn)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                if rep == 1:
                    res.append(f"bn{tn or 'auto'}: {2.0 * M * N * K / ms / 1e9:6.0f}")
        print(f"M={M:6d} {name:15s} N={N:6d} K={K:6d} | " + "  ".join(res), flush=True)
        del Ws
