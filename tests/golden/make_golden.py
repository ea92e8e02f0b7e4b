"""Generates the committed golden fixtures from the ORACLE (oracle/*.py), float64 on CPU, fully seeded.

The reference ships no golden vectors for this path and cannot be imported in this image (diffusers / peft / s4torch
absent), so these fixtures pin the *oracle* (and through it the CUDA path) against regressions; they are not outputs of
the reference itself.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import cs3_dgf as OC  # noqa: E402
from oracle import flux_dit as O  # noqa: E402
from oracle import sampler as OS  # noqa: E402

TINY = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)


def tiny_dit_case():
    """Seeded inputs of the tiny-DiT golden (shared with tests/)."""
    cfg = O.FluxConfig(**TINY)
    P = O.init_params(cfg, seed=1234, dtype=torch.float32, w_std=0.05, bias_std=0.05, lora_b_std=0.05)
    P = {k: v.to(torch.bfloat16) for k, v in P.items()}  # weights as the CUDA path stores them
    g = torch.Generator().manual_seed(2024)
    B, nt, ni, nc = 1, 128, 128, 128
    inp = dict(
        lat=torch.randn(B, ni, 64, generator=g).bfloat16(), cond=torch.randn(B, nc, 64, generator=g).bfloat16(),
        pe=(torch.randn(B, nt, 256, generator=g) * 0.5).bfloat16(), pooled=torch.randn(B, 64, generator=g).bfloat16(),
        img_ids=OS.prepare_latent_image_ids(16, 32), txt_ids=torch.zeros(nt, 3), t=0.65, guidance=3.5)
    inp["cond_ids"] = OS.condition_ids(inp["img_ids"], [0, -16])
    return cfg, P, inp


def tiny_dit_forward(cfg, P, inp, dtype):
    c = lambda x: x.to(dtype)  # noqa: E731
    Pd = {k: v.to(dtype) for k, v in P.items()}
    B = inp["lat"].shape[0]
    return O.tranformer_forward(Pd, cfg, c(inp["cond"]), inp["cond_ids"], None, {}, 0, hidden_states=c(inp["lat"]),
                                encoder_hidden_states=c(inp["pe"]), pooled_projections=c(inp["pooled"]),
                                timestep=torch.full((B,), inp["t"], dtype=dtype), img_ids=inp["img_ids"],
                                txt_ids=inp["txt_ids"], guidance=torch.full((B,), inp["guidance"], dtype=dtype))


def small_cs3_case():
    torch.manual_seed(77)
    s4 = OC.S4Model(4, 8, 8, 2, 8, 96).eval()
    duan = OC.DUAN(16).eval()
    g = torch.Generator().manual_seed(78)
    u = torch.randn(2, 96, 4, generator=g)
    x = torch.randn(2, 16, 40, generator=g)
    c = torch.randn(2, 16, 40, generator=g)
    return s4, duan, u, x, c


def main():
    cfg, P, inp = tiny_dit_case()
    with torch.no_grad():
        out64 = tiny_dit_forward(cfg, P, inp, torch.float64)
    s4, duan, u, x, c = small_cs3_case()
    with torch.no_grad():
        y_s4 = s4(u)
        y_duan, imp, mask = duan(x, c, return_aux=True)
    sig = {f"sigmas_n{n}_L{L}": OS.flow_match_sigmas(n, L).numpy() for n in (4, 28, 50) for L in (1024, 4096)}
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), dit_out64=out64.numpy(), s4_out=y_s4.numpy(),
                        duan_out=y_duan.numpy(), duan_mask=mask.numpy(), **sig)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), "dit std", float(out64.std()))


if __name__ == "__main__":
    main()
