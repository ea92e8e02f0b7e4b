"""CS3 / DGF training-step probe (development aid + ncu target): conditioning forward in training mode + its backward at
per-GPU batch B, without the DiT in between (random upstream gradients)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from loongx_b200 import cs3_bwd as CB
from loongx_b200.config import FluxConfig
from src.train.model import OminiModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = "cuda"
kw = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=4096, pooled_projection_dim=768)
m = OminiModel(FluxConfig(**kw), lora_config={"r": 4, "lora_alpha": 4}, device=dev, model_config={}, fuse_flag=True)
g = torch.Generator().manual_seed(3)
r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(dev)  # noqa: E731
pe, po = r(B, 512, 4096, scale=0.3), r(B, 768)
sig = dict(eeg=r(B, 4, 5000), fnirs=r(B, 6, 600), ppg=r(B, 4, 256), motion=r(B, 6, 100))
params = CB.trainable_parameters(m)
flat = torch.zeros(CB.grad_elements(params), device=dev)
views = CB.grad_views(params, flat)
g1, g2 = torch.randn(B, 512, 4096, device=dev) / 1e3, torch.randn(B, 768, device=dev) / 30
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(4):
    e[0].record()
    _, _, ctx = CB.step_conditioning_train(m, pe, po, sig["eeg"], sig["fnirs"], sig["ppg"], sig["motion"], training=True, seed=it)
    e[1].record()
    flat.zero_()
    CB.step_conditioning_backward(m, ctx, g1, g2, views)
    e[2].record()
    torch.cuda.synchronize()
print(f"B={B}: conditioning forward {e[0].elapsed_time(e[1]):.2f} ms, backward {e[1].elapsed_time(e[2]):.2f} ms")
