"""Is the vendor GEMM any faster than lx_gemm_bf16 once the power cap, not the burst clock, sets the pace?

`scripts/gemm_vs_cublas.py` compares the two in bursts of a few milliseconds (SM clock 1.9 GHz).  Inside the denoise loop the
GPU sits at its power cap (1.40 GHz), where the cost of a kernel is its energy.  This probe runs the GEMM mix of one denoise
step (19 x [qkv, attn_out, ff_up, ff_down] + 38 x [single qkv|mlp, single out], M = 2560, rotating weight copies larger
than the L2) for several seconds per library, alternating, and reports sustained TFLOP/s and the SM clock under load.
Development aid, not a bench value.

  python scripts/gemm_sustained.py [passes_per_leg] [rounds] [legs, e.g. lx,cublas,lx256,lx224,lx192]
  (lxNNN: lx_gemm_bf16 with the N tile forced to NNN instead of the wave-count heuristic of gemm.cu::pick_tile_n)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from loongx_b200 import ops

passes = int(sys.argv[1]) if len(sys.argv) > 1 else 160
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
M = 2560
shapes = {"qkv": (9216, 3072), "attn_out": (3072, 3072), "ff_up": (12288, 3072), "ff_down": (3072, 12288),
          "single_qkv_mlp": (21504, 3072), "single_out": (3072, 15360)}
NCOPY = 4
W = {k: [torch.randn(n, kk, device="cuda", dtype=torch.bfloat16) * 0.02 for _ in range(NCOPY)] for k, (n, kk) in shapes.items()}
A = {kk: torch.randn(M, kk, device="cuda", dtype=torch.bfloat16) for kk in {v[1] for v in shapes.values()}}
O = {n: torch.empty(M, n, device="cuda", dtype=torch.bfloat16) for n in {v[0] for v in shapes.values()}}
seq = ["qkv", "attn_out", "ff_up", "ff_down"] * 19 + ["single_qkv_mlp", "single_out"] * 38
flop_pass = sum(2.0 * M * shapes[s][0] * shapes[s][1] for s in seq)


def one_pass(which, it):
    for j, s in enumerate(seq):
        n, kk = shapes[s]
        w = W[s][(it + j) % NCOPY]
        if which.startswith("lx"):
            ops.gemm(A[kk], w, None, O[n], tile_n=int(which[2:] or 0))
        else:
            torch.matmul(A[kk], w.t(), out=O[n])


try:  # in-process NVML: a forked nvidia-smi per sample stalls the launching thread long enough to drain the GPU
    import pynvml

    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
except Exception:  # noqa: BLE001
    _h = None


def clock():
    if _h is None:
        return None
    return (pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(_h) / 1e3)


print(f"GEMM mix of one denoise step: {flop_pass / 1e12:.2f} TFLOP per pass, {len(seq)} launches", flush=True)
for r in range(rounds):
    for which in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("lx", "cublas")):
        for it in range(3):
            one_pass(which, it)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk = []
        e0.record()
        for n in range(passes):
            one_pass(which, n)
            if n % 8 == 7:  # (the launch queue is bounded, so the host stays ~1000 launches = 7 passes ahead of the device)
                clk.append(clock())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        clk = [c for c in clk[len(clk) // 3:] if c is not None]
        mhz = sorted(c[0] for c in clk)[len(clk) // 2] if clk else None
        watt = sorted(c[1] for c in clk)[len(clk) // 2] if clk else None
        print(f"round {r} {which:6s}: {passes} passes in {ms / 1e3:.2f} s -> {passes * flop_pass / ms / 1e9:7.1f} TFLOP/s sustained, "
              f"{ms / passes:.2f} ms per pass; median SM clock {mhz} MHz, power {watt} W", flush=True)
