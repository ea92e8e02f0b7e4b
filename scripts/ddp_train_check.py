"""2+ GPU check of the training collective (SURVEY.md §8e "Training collective"): one process per GPU under torchrun,
identical replicas from the same seed, different batches per rank; OminiModel.step(batch).backward() must leave every
rank with the MEAN of the per-rank LoRA gradients (one NCCL all-reduce over the flat bucket), bit-identical across ranks.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 scripts/ddp_train_check.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from loongx_b200.config import FluxConfig
from src.train.model import OminiModel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
BRAIN = bool(os.environ.get("LX_DDP_BRAIN"))  # with the CS3 / DGF conditioning: the bucket carries the encoder gradients too
jd, pd, nt = (4096, 768, 512) if BRAIN else (256, 64, 128)
cfg = FluxConfig(num_layers=2, num_single_layers=2, num_attention_heads=2, joint_attention_dim=jd, pooled_projection_dim=pd)
m = OminiModel(cfg, lora_config={"r": 4, "lora_alpha": 4}, device=str(dev), model_config={}, use_brain_condition=BRAIN, seed=7)
g = torch.Generator().manual_seed(100 + rank)  # a different batch on every rank
B, h, w = 2, 16, 32
r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).bfloat16().to(dev)  # noqa: E731
batch = dict(image=r(B, 16, h, w), condition=r(B, 16, h, w), prompt_embeds=r(B, nt, jd, scale=0.5),
             pooled_prompt_embeds=r(B, pd), position_delta=[[0, -16]], condition_type=["subject"] * B,
             t=torch.tensor([0.3, 0.7]), noise=r(B, 128, 64))
if BRAIN:
    batch.update(eeg=r(B, 4, 5000).float(), fnirs=r(B, 6, 600).float(), ppg=r(B, 4, 256).float(), motion=r(B, 6, 100).float())
loss = m.step(batch)
tr = m._trainer_obj
# rank-local gradients, no collective
tr.zero_grad()
tr.backward(1.0)
if BRAIN:
    from loongx_b200 import cs3_bwd as CB
    from loongx_b200.train import EncoderBackward

    EncoderBackward(m, loss.grad_fn.enc.ctx, CB.trainable_parameters(m), tr).backward(tr.d_prompt, tr.d_pooled)
local_grad = tr.grad_flat.clone()
gathered = [torch.zeros_like(local_grad) for _ in range(world)]
dist.all_gather(gathered, local_grad)
mean = torch.stack(gathered).mean(0)
# the product path: autograd node -> native backward -> ONE all-reduce(mean) of the flat bucket
loss.backward()
if BRAIN:  # what autograd left on the parameters, laid out like the flat bucket (16-byte aligned encoder slices)
    enc_params = CB.trainable_parameters(m)
    layout, total = CB.grad_layout(enc_params)
    enc_flat = torch.zeros(total, device=dev)
    for p, (o, n) in zip(enc_params, layout):
        if p.grad is not None:
            enc_flat[o:o + n] = (torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).flatten()
    got = torch.cat([p.grad.flatten() for p in tr.parameters()] + [enc_flat])
else:
    got = torch.cat([p.grad.flatten() for p in tr.parameters()])
err = ((got - mean).norm() / mean.norm()).item()
all_got = [torch.zeros_like(got) for _ in range(world)]
dist.all_gather(all_got, got)
same = all(torch.equal(all_got[0], x) for x in all_got)
differs = (gathered[0] - gathered[-1]).abs().max().item() > 0
if rank == 0:
    print(f"world {world}: loss[rank0] {float(loss):.5f}; |all-reduced - mean(local)| rel {err:.3e}; identical across ranks: {same}; "
          f"local grads differ between ranks: {differs}; bucket {got.numel()} fp32")
# (the encoder backward accumulates with fp32 atomics: two runs of the same local backward differ at the 1e-7 level)
assert err < (1e-5 if BRAIN else 1e-6) and same and differs
dist.barrier()
dist.destroy_process_group()
