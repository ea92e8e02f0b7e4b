"""Device-side timeline of one denoise step (development aid): every GEMM / attention / ln_modulate launch records the
%globaltimer of its first CTA's entry, of the first CTA past its programmatic-dependency wait and of its last CTA's exit.
Prints per-class busy time, the gaps between consecutive kernels and the slowest launches, for the real PDL-chained loop
(nothing is enqueued between kernels, unlike CUDA-event profiling)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from loongx_b200 import _lib as L
from loongx_b200.config import FluxConfig
from loongx_b200.dit import DitPlan, DitWeights, random_params

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
res = int(sys.argv[2]) if len(sys.argv) > 2 else 512
T = 3
cfg = FluxConfig()
dev = "cuda"
W = DitWeights(random_params(cfg, dev), cfg, dev, consume=True)
n = (res // 16) ** 2
nt, ni, nc = 512, n, n
plan = DitPlan(W, B, nt, ni, nc, T=T, model_config={})
h = res // 16


def ids(dc=0):
    i = torch.zeros(h, h, 3)
    i[..., 1] += torch.arange(h)[:, None]
    i[..., 2] += torch.arange(h)[None, :] + dc
    return i.reshape(-1, 3)


plan.set_ids(torch.zeros(nt, 3), ids(), ids(-h))
g = torch.Generator(device=dev).manual_seed(0)
pe = (torch.randn(B, nt, 4096, generator=g, device=dev) * 0.1).bfloat16()
pooled = torch.randn(B, 768, generator=g, device=dev).bfloat16()
cond = torch.randn(B, nc, 64, generator=g, device=dev).bfloat16()
lat = torch.randn(B, ni, 64, generator=g, device=dev).bfloat16()
plan.prepare(pe, pooled, cond, [1.0 - 0.3 * s for s in range(T) for _ in range(B)], [3.5] * B)
out = torch.empty_like(lat)
for rep in range(12):  # warm up: clocks settle under the power cap
    plan.step(rep % T, lat, out)
torch.cuda.synchronize()
cap = 1024
buf = torch.zeros((cap, 4), dtype=torch.int64, device=dev)
buf[:, 0] = buf[:, 1] = 2 ** 62
L.lib.lx_debug_timeline.argtypes = [C.c_void_p, C.c_int]
L.lib.lx_debug_timeline(buf.data_ptr(), cap)
atr = torch.zeros(148 * 16, dtype=torch.int64, device=dev)  # per-CTA trace of the (last) attention launch of the step
L.lib.lx_attention_debug_cta_trace.argtypes = [C.c_void_p]
L.lib.lx_attention_debug_cta_trace(atr.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
plan.step(1, lat, out)
e1.record()
torch.cuda.synchronize()
cnt = L.lib.lx_debug_timeline_count()
cls = [L.lib.lx_debug_timeline_class(i) for i in range(cnt)]
L.lib.lx_debug_timeline(None, 0)
L.lib.lx_attention_debug_cta_trace(None)
at = [r for r in atr.cpu().view(148, 16).tolist() if r[1] != 0]
if at:
    import statistics as st
    mhz = [(r[5] - r[1]) / max(r[15] - r[14], 1) * 1e3 for r in at]
    print(f"in-loop attention (last launch): {len(at)} CTAs, SM clock {st.median(mhz):.0f} MHz, per-CTA total clk median "
          f"{int(st.median([r[5] - r[1] for r in at]))} max {max(r[5] - r[1] for r in at)}, setup {int(st.median([r[2] - r[1] for r in at]))}, "
          f"first S {int(st.median([r[3] - r[2] for r in at]))}, flag wait max {max(r[6] for r in at)}, wall "
          f"{(max(r[15] for r in at) - min(r[14] for r in at)) / 1e3:.1f} us")
rows = buf[:cnt].cpu().tolist()
t0 = rows[0][0]
names = {0: "gemm", 1: "attn", 2: "ln"}
print(f"step: {e0.elapsed_time(e1):.3f} ms by CUDA events; {cnt} launches; first entry -> last exit {(rows[-1][2] - t0) / 1e6:.3f} ms")
busy = {0: 0, 1: 0, 2: 0}
gaps = {}
prev_end, prev_cls = None, None
for c, (a, w, e, _) in zip(cls, rows):
    busy[c] += e - w
    if prev_end is not None:
        k = f"{names[prev_cls]}->{names[c]}"
        gaps.setdefault(k, []).append(w - prev_end)
    prev_end, prev_cls = e, c
tot = rows[-1][2] - rows[0][1]
print("busy (past-wait -> last exit) per class, us:", {names[k]: round(v / 1e3, 1) for k, v in busy.items()},
      " sum", round(sum(busy.values()) / 1e3, 1), " wall", round(tot / 1e3, 1))
for k, v in sorted(gaps.items()):
    v2 = sorted(v)
    print(f"  gap {k:12s} n={len(v):3d}  median {v2[len(v2) // 2] / 1e3:6.2f} us  mean {sum(v) / len(v) / 1e3:6.2f}  max {v2[-1] / 1e3:6.2f}  total {sum(v) / 1e3:8.1f} us")
print("launch -> past-wait (prologue overlapped with the predecessor), median us per class:",
      {names[c]: round(sorted([(w - a) for cc, (a, w, e, _) in zip(cls, rows) if cc == c])[len([1 for cc in cls if cc == c]) // 2] / 1e3, 2)
       for c in (0, 1, 2)})
# one double block and one single block in detail
def show(lo, hi):
    for i in range(lo, hi):
        a, w, e, _ = rows[i]
        print(f"    #{i:3d} {names[cls[i]]:5s} entry {(a - t0) / 1e3:9.2f}  start {(w - t0) / 1e3:9.2f}  end {(e - t0) / 1e3:9.2f}  dur {(e - w) / 1e3:7.2f} us")
print("  double block 5:")
show(1 + 5 * 7, 1 + 6 * 7)
print("  single block 5:")
show(1 + 19 * 7 + 5 * 4, 1 + 19 * 7 + 6 * 4)
