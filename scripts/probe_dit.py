"""Full-size FLUX.1-dev-shaped DiT step timing probe (synthetic weights), prints per-step ms and TFLOP/s."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loongx_b200.config import FluxConfig
from loongx_b200.dit import DitWeights, DitPlan, random_params, euler_step

from loongx_b200 import _lib as _L
_L.lib.lx_debug_set_pdl(int(os.environ.get("LX_PDL", "1")))  # A/B switch for programmatic dependent launch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
res = int(sys.argv[2]) if len(sys.argv) > 2 else 512
T = int(sys.argv[3]) if len(sys.argv) > 3 else 4
_l = os.environ.get("LX_LAYERS")  # e.g. "1,1": shallow model at full width for ncu captures
cfg = FluxConfig(num_layers=int(_l.split(",")[0]), num_single_layers=int(_l.split(",")[1])) if _l else FluxConfig()
dev = "cuda"
t0 = time.time()
P = random_params(cfg, dev)
W = DitWeights(P, cfg, dev, consume=True)
del P
torch.cuda.synchronize()
print(f"weights: {W.param_bytes()/1e9:.2f} GB packed in {time.time()-t0:.1f}s, mem {torch.cuda.memory_allocated()/1e9:.1f} GB")
n = (res // 16) ** 2
nt, ni, nc = 512, n, n
_mc = {"independent_condition": True} if os.environ.get("LX_INDEP") else {}
plan = DitPlan(W, B, nt, ni, nc, T=T, model_config=_mc, cache_cond=bool(os.environ.get("LX_CACHE_COND")))
print("model_config", _mc, "cache_cond", plan.cache_cond)
h = res // 16
def ids(dc=0):
    i = torch.zeros(h, h, 3); i[..., 1] += torch.arange(h)[:, None]; i[..., 2] += torch.arange(h)[None, :] + dc
    return i.reshape(-1, 3)
plan.set_ids(torch.zeros(nt, 3), ids(), ids(-h))
g = torch.Generator(device=dev).manual_seed(0)
pe = (torch.randn(B, nt, 4096, generator=g, device=dev) * 0.1).bfloat16()
pooled = torch.randn(B, 768, generator=g, device=dev).bfloat16()
cond = torch.randn(B, nc, 64, generator=g, device=dev).bfloat16()
lat = torch.randn(B, ni, 64, generator=g, device=dev).bfloat16()
ts = [1.0 - 0.9 * s / T for s in range(T) for _ in range(B)]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); plan.prepare(pe, pooled, cond, ts, [3.5] * B); e1.record(); torch.cuda.synchronize()
print(f"prepare (T={T}): {e0.elapsed_time(e1):.2f} ms (first call)")
import time as _t
_h0 = _t.perf_counter()
e0.record(); plan.prepare(pe, pooled, cond, ts, [3.5] * B); e1.record(); torch.cuda.synchronize()
print(f"prepare (T={T}): {e0.elapsed_time(e1):.2f} ms device, {(_t.perf_counter() - _h0) * 1e3:.2f} ms host wall (second call)")
out = torch.empty_like(lat)
for s in range(min(2, T)):
    plan.step(s, lat, out)
torch.cuda.synchronize()
print("finite:", torch.isfinite(out.float()).all().item(), "std", out.float().std().item())
S = nt + ni + nc
D = 3072
flop = B * (57 * (24 * D * D * S + 4 * S * S * D) + 2 * 64 * D * (ni + nc) + 2 * 4096 * D * nt + 2 * D * 64 * ni)
t0 = time.time()
e0.record()
for s in range(T):
    plan.step(s, lat, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / T
print(f"B={B} res={res}: {ms:.2f} ms/step  {flop/ms/1e9:.1f} TFLOP/s (algorithmic)  host wall {(time.time()-t0)/T*1e3:.2f} ms/step")
# marginal cost of a kernel class inside the real loop (lx_debug_skip: the launcher returns early, buffers keep their last
# realistic contents so the data-dependent power draw of the other kernels does not change)
if os.environ.get("LX_MARGINAL"):
    for name, mask in (("ln_modulate", 4), ("attention", 2)):
        _L.lib.lx_debug_skip(mask)
        for s in range(min(2, T)):
            plan.step(s, lat, out)
        e0.record()
        for s in range(T):
            plan.step(s, lat, out)
        e1.record(); torch.cuda.synchronize()
        print(f"  without {name}: {e0.elapsed_time(e1) / T:.2f} ms/step  (marginal cost {ms - e0.elapsed_time(e1) / T:.2f} ms/step)")
    _L.lib.lx_debug_skip(0)
