"""AdaLN row kernel of the shipped library (lx_ln_modulate) on the edit's [2560, 3072] activation: average launch time
over back-to-back launches (inputs L2-resident, as inside the denoise loop) and over a rotating set of buffers larger than
the L2 (DRAM-streaming).  Development aid, not a bench value.

  python scripts/ln_lib_probe.py [rows] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from loongx_b200 import ops
from loongx_b200.train import ln_modulate

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2560
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
D, dev = 3072, "cuda"
B = max(1, rows // 2560)
tm = ops.make_tile_meta(B, 512, 1024, 1024, dev) if rows == B * 2560 else ops.make_tile_meta(1, rows, 0, 0, dev)
g = torch.Generator(device=dev).manual_seed(0)
mod = [(torch.randn(B, D, generator=g, device=dev) * 0.1).bfloat16() for _ in range(2)]
nbuf = 12  # 12 x 2 x 15.7 MB = 377 MB >> 126 MB L2
xs = [torch.randn(rows, D, generator=g, device=dev).bfloat16() for _ in range(nbuf)]
outs = [torch.empty_like(x) for x in xs]


def run(k):
    ln_modulate(xs[k], outs[k], tm, [mod[0]] * 3, [mod[1]] * 3)


for label, pick in (("L2-resident (same buffers)", lambda i: 0), ("DRAM-streaming (rotating 377 MB)", lambda i: i % nbuf)):
    for i in range(20):
        run(pick(i))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        run(pick(i))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"ln_modulate [{rows}, {D}] {label}: {us:.2f} us / launch, {2 * rows * D * 2 / us / 1e3:.0f} GB/s")
print("finite:", bool(torch.isfinite(outs[0].float()).all()), "| rows", rows, "| lib", os.environ.get("LX_LIB", "in-tree"))
