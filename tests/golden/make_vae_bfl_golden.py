"""Pins oracle/vae.py against an independent, executable implementation of the FLUX.1 autoencoder: Black Forest Labs'
`AutoEncoder` as shipped in the `torchtitan` package of this image (torchtitan/experiments/flux/model/autoencoder.py, the
architecture diffusers' AutoencoderKL loads for FLUX.1-dev: ch 128, ch_mult (1, 2, 4, 4), 2 resnets per level, z 16,
scale 0.3611, shift 0.1159).  The oracle keeps its parameters in diffusers' state-dict naming; `to_bfl_state_dict` is the
published key correspondence between the two layouts (diffusers' convert_ldm_vae_checkpoint run backwards).

  python tests/golden/make_vae_bfl_golden.py        # writes tests/golden/vae_bfl_v1.npz (reduced width, seeded)

tests/test_vae_cpu.py replays the fixture (always) and re-runs the BFL module live (where torchtitan imports).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vae as V  # noqa: E402

FIXTURE = os.path.join(ROOT, "tests", "golden", "vae_bfl_v1.npz")
SMALL = dict(block_out_channels=(32, 64, 128, 128))  # the fixture's width: GroupNorm(32) needs multiples of 32


def to_bfl_state_dict(P, cfg):
    """oracle / diffusers names -> the BFL module's state dict (1x1 convolutions for the attention projections)."""
    n = len(cfg.block_out_channels)
    out = {}

    def put(dst, src, conv1x1=False):
        for s in ("weight", "bias"):
            t = P[f"{src}.{s}"]
            out[f"{dst}.{s}"] = t[:, :, None, None].clone() if (conv1x1 and s == "weight") else t.clone()

    def resnet(dst, src):
        for k in ("norm1", "conv1", "norm2", "conv2"):
            put(f"{dst}.{k}", f"{src}.{k}")
        if f"{src}.conv_shortcut.weight" in P:
            put(f"{dst}.nin_shortcut", f"{src}.conv_shortcut")

    def mid(dst, src):
        resnet(f"{dst}.block_1", f"{src}.resnets.0")
        resnet(f"{dst}.block_2", f"{src}.resnets.1")
        put(f"{dst}.attn_1.norm", f"{src}.attentions.0.group_norm")
        for a, b in (("q", "to_q"), ("k", "to_k"), ("v", "to_v"), ("proj_out", "to_out.0")):
            put(f"{dst}.attn_1.{a}", f"{src}.attentions.0.{b}", conv1x1=True)

    put("encoder.conv_in", "encoder.conv_in")
    for i in range(n):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down.{i}.block.{j}", f"encoder.down_blocks.{i}.resnets.{j}")
        if i != n - 1:
            put(f"encoder.down.{i}.downsample.conv", f"encoder.down_blocks.{i}.downsamplers.0.conv")
    mid("encoder.mid", "encoder.mid_block")
    put("encoder.norm_out", "encoder.conv_norm_out")
    put("encoder.conv_out", "encoder.conv_out")
    put("decoder.conv_in", "decoder.conv_in")
    mid("decoder.mid", "decoder.mid_block")
    for i in range(n):  # diffusers' up_blocks run lowest resolution first; the BFL list is indexed by resolution level
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up.{n - 1 - i}.block.{j}", f"decoder.up_blocks.{i}.resnets.{j}")
        if i != n - 1:
            put(f"decoder.up.{n - 1 - i}.upsample.conv", f"decoder.up_blocks.{i}.upsamplers.0.conv")
    put("decoder.norm_out", "decoder.conv_norm_out")
    put("decoder.conv_out", "decoder.conv_out")
    return out


def bfl_autoencoder(P, cfg):
    from torchtitan.experiments.flux.model.autoencoder import AutoEncoder, AutoEncoderParams

    ch = cfg.block_out_channels[0]
    ae = AutoEncoder(AutoEncoderParams(resolution=64, in_channels=cfg.in_channels, ch=ch, out_ch=cfg.out_channels,
                                       ch_mult=tuple(c // ch for c in cfg.block_out_channels),
                                       num_res_blocks=cfg.layers_per_block, z_channels=cfg.latent_channels,
                                       scale_factor=cfg.scaling_factor, shift_factor=cfg.shift_factor)).float().eval()
    missing, unexpected = ae.load_state_dict(to_bfl_state_dict(P, cfg), strict=True), None
    return ae


def run_bfl(P, cfg, images, latents, seed):
    """(moments, encode(images) with the module's own sampling under `seed`, decode(latents)) from the BFL module."""
    ae = bfl_autoencoder(P, cfg)
    with torch.no_grad():
        moments = ae.encoder(images)
        torch.manual_seed(seed)
        z = ae.encode(images)
        img = ae.decode(latents)
    return moments, z, img


def inputs(seed=11, hw=(40, 56)):
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(2, 3, hw[0], hw[1], generator=g) * 2 - 1
    latents = torch.randn(2, 16, hw[0] // 8, hw[1] // 8, generator=g) * 0.8
    return images, latents


if __name__ == "__main__":
    cfg = V.VaeConfig(**SMALL)
    P = V.init_params(cfg, seed=77)
    images, latents = inputs()
    moments, z, img = run_bfl(P, cfg, images, latents, seed=5)
    np.savez_compressed(FIXTURE, images=images.numpy(), latents=latents.numpy(), moments=moments.numpy(), z=z.numpy(),
                        img=img.numpy())
    print(FIXTURE, {k: tuple(v.shape) for k, v in dict(moments=moments, z=z, img=img).items()})
