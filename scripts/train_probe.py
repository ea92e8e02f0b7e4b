"""Full-size (FLUX.1-dev geometry) training-step probe: OminiModel.step forward + backward at 512x512 with an image
condition (S = 512 + 1024 + 1024), LoRA r=4, per-GPU batch B.  Prints ms / step, algorithmic TFLOP/s (SURVEY.md §8d:
fwd F + recompute F + bwd [linear dX 57*24*D^2*S + attention 2.5 x]) and peak memory.  Synthetic weights / inputs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from loongx_b200.config import FluxConfig
from loongx_b200.dit import DitWeights, random_params
from loongx_b200.train import DitTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
layers = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (19, 38)
dev = "cuda"
cfg = FluxConfig(num_layers=layers[0], num_single_layers=layers[1])
t0 = time.time()
W = DitWeights(random_params(cfg, dev), cfg, dev, consume=True)
nt, ni, nc = 512, 1024, 1024
_rc = os.environ.get("LX_RECOMPUTE")
tr = DitTrainer(W, B, nt, ni, nc, model_config={}, recompute=None if _rc is None else bool(int(_rc)))
print("recompute", tr.recompute)
torch.cuda.synchronize()
print(f"weights + transposed panels + workspace: {torch.cuda.memory_allocated() / 1e9:.1f} GB in {time.time() - t0:.1f} s; "
      f"{len(tr.factors)} LoRA targets, {tr.grad_flat.numel() / 1e6:.2f} M trainable", flush=True)
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s, scale=1.0: (torch.randn(*s, generator=g, device=dev) * scale).bfloat16()  # noqa: E731
x0, x1, cond = r(B, ni, 64), r(B, ni, 64), r(B, nc, 64)
pe, po = r(B, nt, 4096, scale=0.1), r(B, 768)
t = torch.sigmoid(torch.randn(B, generator=g, device=dev))
side = 32
ids = torch.zeros(side, side, 3)
ids[..., 1] += torch.arange(side)[:, None]
ids[..., 2] += torch.arange(side)[None, :]
ids = ids.reshape(-1, 3).to(dev)
cids = ids.clone()
cids[:, 2] -= side
txt = torch.zeros(nt, 3, device=dev)
D, S = cfg.inner_dim, nt + ni + nc
nb = layers[0] + layers[1]
F = nb * (24 * D * D * S + 4 * S * S * D)
flop = B * (2 * F + nb * 24 * D * D * S + 2.5 * nb * 4 * S * S * D)
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(steps + 1):
    e[0].record()
    loss = tr.forward(x0, x1, t, cond, pe, po, txt, ids, cids, 1.0)
    e[1].record()
    tr.zero_grad()
    tr.backward()
    e[2].record()
    torch.cuda.synchronize()
    gn = float(tr.grad_flat.norm())
    f_ms, b_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    print(f"step {it}: loss {loss.item():.5f} |grad| {gn:.4e} fwd {f_ms:.1f} ms bwd(+recompute) {b_ms:.1f} ms "
          f"-> {flop / (f_ms + b_ms) / 1e9:.0f} algorithmic TFLOP/s, peak mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB",
          flush=True)
    assert torch.isfinite(tr.grad_flat).all() and gn > 0
