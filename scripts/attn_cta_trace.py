"""Per-CTA timeline of the persistent attention kernel (development aid): lx_attention_debug_cta_trace fills 16 slots
per CTA; prints the split-work schedule's segment / epilogue / flag-wait times and a split on / off A-B timing."""
import ctypes as C
import os
import statistics as st
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from loongx_b200 import _lib as L
from loongx_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H, nt, ni, nc = 24, 512, 1024, 1024
if len(sys.argv) > 2:
    ni = nc = int(sys.argv[2])
S = nt + ni + nc
g = torch.Generator(device="cuda").manual_seed(1)
q, k, v = [torch.randn((B, H, S, 128), generator=g, device="cuda").bfloat16() for _ in range(3)]
out = torch.empty((B * S, H * 128), device="cuda", dtype=torch.bfloat16)
orb = ops.make_out_row_base(B, nt, ni, nc, "cuda")
L.lib.lx_attention_debug_cta_trace.argtypes = [C.c_void_p]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for split in (0, 1):
    L.lib.lx_debug_attention_split(split)
    for _ in range(3):
        ops.attention(q, k, v, out, orb, n_cond=nc)
    e0.record()
    for _ in range(20):
        ops.attention(q, k, v, out, orb, n_cond=nc)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"split={split}: {ms * 1e3:.1f} us  {4 * B * H * S * S * 128 / ms / 1e9:.0f} TFLOP/s")
    n = 148
    tr = torch.zeros(n * 16, dtype=torch.int64, device="cuda")
    L.lib.lx_attention_debug_cta_trace(tr.data_ptr())
    ops.attention(q, k, v, out, orb, n_cond=nc)
    torch.cuda.synchronize()
    L.lib.lx_attention_debug_cta_trace(None)
    t = [r for r in tr.cpu().view(n, 16).tolist() if r[1] != 0]
    med = lambda xs: int(st.median(xs)) if xs else 0  # noqa: E731
    print(f"  CTAs {len(t)}  clk: setup {med([r[2] - r[1] for r in t])}  first S {med([r[3] - r[2] for r in t])}  "
          f"entry->last loop end {med([r[4] - r[1] for r in t])}  last epilogue+exit {med([r[5] - r[4] for r in t])}  "
          f"total med {med([r[5] - r[1] for r in t])} max {max(r[5] - r[1] for r in t)} min {min(r[5] - r[1] for r in t)}")
    print(f"  segments per CTA: {sorted(set(r[7] for r in t))}  flag wait clk: med {med([r[6] for r in t])} max {max(r[6] for r in t)}")
    for si in range(3):
        rows = [r for r in t if r[7] > si]
        if rows:
            print(f"  segment {si}: n={len(rows)}  loop end at {med([r[8 + 2 * si] - r[1] for r in rows])} (since entry)  "
                  f"epilogue {med([r[9 + 2 * si] - r[8 + 2 * si] for r in rows])} clk (max {max(r[9 + 2 * si] - r[8 + 2 * si] for r in rows)})")
    g0 = min(r[14] for r in t)
    print(f"  globaltimer ns: entry spread {max(r[14] for r in t) - g0}  exit: first {min(r[15] for r in t) - g0} last {max(r[15] for r in t) - g0}")
L.lib.lx_debug_attention_split(1)
