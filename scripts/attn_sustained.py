"""Attention of one edit (B = 1, 24 heads, S = 2560, head dim 128, no mask) SUSTAINED for seconds: lx_attention against the
vendor kernels torch ships (scaled_dot_product_attention through the cuDNN and the flash back ends), alternating, with
SM clock and board power sampled in-process.  Answers two questions the bursts of a few launches cannot: is the attention
kernel itself power-capped, and how far is it from the vendor's Blackwell kernel on this shape?  Development aid.

  python scripts/attn_sustained.py [launches_per_leg] [rounds] [B] [legs, e.g. lx,cudnn]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

from loongx_b200 import _lib as L
from loongx_b200 import ops

n_launch = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
H, nt, ni, nc = 24, 512, 1024, 1024
S = nt + ni + nc
g = torch.Generator(device="cuda").manual_seed(1)
NSET = 6  # rotating operand sets (6 x 47 MB: K / V come from the L2 or HBM as in the loop, never from one hot set)
qkv = [[torch.randn((B, H, S, 128), generator=g, device="cuda").bfloat16() for _ in range(3)] for _ in range(NSET)]
out = torch.empty((B * S, H * 128), device="cuda", dtype=torch.bfloat16)
orb = ops.make_out_row_base(B, nt, ni, nc, "cuda")
flop = 4.0 * B * H * S * S * 128
if os.environ.get("LX_ATT_SPLIT"):  # A/B: 0 = units are never cut between CTAs (no partial exchange, 1.62 waves)
    L.lib.lx_debug_attention_split(int(os.environ["LX_ATT_SPLIT"]))

try:
    import pynvml

    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
except Exception:  # noqa: BLE001
    _h = None


def sample():
    if _h is None:
        return None
    return (pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(_h) / 1e3)


def lx(i):
    q, k, v = qkv[i % NSET]
    ops.attention(q, k, v, out, orb, n_cond=nc)


def sdpa(backend):
    def f(i):
        q, k, v = qkv[i % NSET]
        with sdpa_kernel(backend):
            F.scaled_dot_product_attention(q, k, v)
    return f


legs = [("lx", lx), ("cudnn", sdpa(SDPBackend.CUDNN_ATTENTION)), ("flash", sdpa(SDPBackend.FLASH_ATTENTION))]
if len(sys.argv) > 4:
    legs = [l for l in legs if l[0] in sys.argv[4].split(",")]
ref = F.scaled_dot_product_attention(qkv[0][0].float(), qkv[0][1].float(), qkv[0][2].float())
lx(0)
got = out.view(B, S, H, 128).permute(0, 2, 1, 3).float()
# (row order of `out` is stream-major [txt | img | cond] x batch; at B = 1 it is the token order)
if os.environ.get("LX_ATT_DUMP"):  # A/B of two builds: the outputs must be bit-identical when only the schedule changed
    torch.save(out.cpu(), os.environ["LX_ATT_DUMP"])
if B == 1:
    print(f"lx vs fp32 SDPA relL2 {float((got - ref).norm() / ref.norm()):.3e}", flush=True)
for r in range(rounds):
    for name, fn in legs:
        try:
            for i in range(50):
                fn(i)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"round {r} {name}: unavailable ({type(e).__name__}: {str(e)[:120]})", flush=True)
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        smp = []
        e0.record()
        for i in range(n_launch):
            fn(i)
            if i % 500 == 499:
                smp.append(sample())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        smp = [s for s in smp[len(smp) // 3:] if s is not None]
        mhz = sorted(s[0] for s in smp)[len(smp) // 2] if smp else None
        watt = sorted(s[1] for s in smp)[len(smp) // 2] if smp else None
        print(f"round {r} {name:6s}: {n_launch} launches in {ms / 1e3:.2f} s -> {ms / n_launch * 1e3:7.1f} us / launch, "
              f"{n_launch * flop / ms / 1e9:7.1f} TFLOP/s sustained; median SM clock {mhz} MHz, power {watt} W", flush=True)
