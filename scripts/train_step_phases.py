"""Where the time of one BASELINE configs[4] training step goes (bench.py's c4 leg: OminiModel.step + backward + AdamW at
per-GPU batch 8): CUDA-event and host-clock spans around the phases of the step.  Development aid, not a bench value.

  python scripts/train_step_phases.py [batch] [steps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import loongx_b200.cs3_bwd as CB
import loongx_b200.train as T
from src.train.model import OminiModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
RES, N_TXT = 512, 512
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
SPANS = []


def span(owner, name, label=None):
    fn = getattr(owner, name)
    label = label or name

    def wrapped(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0 = time.perf_counter()
        e0.record()
        out = fn(*a, **k)
        e1.record()
        SPANS.append((label, e0, e1, time.perf_counter() - h0))
        return out

    setattr(owner, name, wrapped)


model = OminiModel("synthetic", lora_config={"r": 4, "lora_alpha": 4}, device=str(dev), model_config={
    "union_cond_attn": True, "add_cond_attn": False, "latent_lora": False}, use_brain_condition=True, fuse_flag=True)
side = RES // 8
g = torch.Generator().manual_seed(7)
r = lambda *s, scale=1.0, dt=torch.bfloat16: (torch.randn(*s, generator=g) * scale).to(dt).to(dev)  # noqa: E731
batch = dict(image=r(B, 16, side, side), condition=r(B, 16, side, side), prompt_embeds=r(B, N_TXT, 4096, scale=0.1),
             pooled_prompt_embeds=r(B, 768), position_delta=[[0, -(RES // 16)]], condition_type=["subject"] * B,
             eeg=r(B, 4, 5000, dt=torch.float32), fnirs=r(B, 6, 600, dt=torch.float32),
             ppg=r(B, 4, 256, dt=torch.float32), motion=r(B, 6, 100, dt=torch.float32))
opt = torch.optim.AdamW(model.lora_layers, lr=1e-4)  # model.py:533-558
span(CB, "step_conditioning_train", "cs3_dgf_forward")
span(T.DitTrainer, "forward", "dit_forward")
span(T.DitTrainer, "backward", "dit_backward")
span(T.DitTrainer, "remerge_if_stale", "remerge_if_stale")
span(T.DitTrainer, "zero_grad", "bucket_zero")
span(T.EncoderBackward, "backward", "cs3_dgf_backward")
span(model, "step", "model.step (whole)")
span(opt, "step", "optimizer.step")


def one_step():
    opt.zero_grad(set_to_none=True)
    loss = model.step(batch)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    e0.record()
    loss.backward()
    e1.record()
    SPANS.append(("loss.backward (whole)", e0, e1, time.perf_counter() - h0))
    opt.step()
    return loss


one_step()
one_step()
torch.cuda.synchronize()
for it in range(steps):
    SPANS.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    e0.record()
    one_step()
    e1.record()
    host = time.perf_counter() - h0
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1)
    print(f"step {it}: {total:.1f} ms on the device, host enqueue {host * 1e3:.1f} ms, micro-batch {model._trainer_obj.B}")
    agg = {}
    for label, a, b, h in SPANS:
        d = agg.setdefault(label, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += a.elapsed_time(b)
        d[2] += h * 1e3
    for label, (n, ms, hms) in agg.items():
        print(f"  {label:28s} x{n}: {ms:8.2f} ms device ({ms / total * 100:5.1f} %), host {hms:7.2f} ms")
print(f"peak memory {torch.cuda.max_memory_allocated() / 1e9:.1f} GB")
