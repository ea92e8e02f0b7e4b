"""ORACLE (test infrastructure only) — CPU restatement of the 16-channel FLUX VAE either side of the denoising loop
(SURVEY.md §8f.2).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file.

Reference call sites:  pipeline_tools.py:7-12  `vae.encode(images).latent_dist.sample()` then `(z - shift) * scale`;
                       generate.py:375-380    `z / scale + shift` -> `vae.decode(z)[0]` -> `image_processor.postprocess`.

The arithmetic lives in a THIRD-PARTY package that is absent from /root/reference and from this image:
diffusers==0.31.0 (train/requirements.txt:1), `AutoencoderKL` with FLUX.1-dev's vae/config.json
(block_out_channels (128, 256, 512, 512), layers_per_block 2, latent_channels 16, norm_num_groups 32, act_fn silu,
scaling_factor 0.3611, shift_factor 0.1159, use_quant_conv / use_post_quant_conv false, mid_block_add_attention true).
This file restates its published algorithm:
  models/autoencoders/vae.py            Encoder.forward / Decoder.forward / DiagonalGaussianDistribution
  models/unets/unet_2d_blocks.py        DownEncoderBlock2D, UpDecoderBlock2D, UNetMidBlock2D (one attention, two resnets)
  models/resnet.py                      ResnetBlock2D (temb = None, eps 1e-6, output_scale_factor 1, 1x1 conv_shortcut when
                                        the channel count changes)
  models/downsampling.py                Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then 3x3 stride-2 convolution
  models/upsampling.py                  Upsample2D: nearest x2 then 3x3 convolution
  models/attention_processor.py         Attention(heads=1, dim_head=C, group_norm 32, residual_connection=True) over the
                                        H*W positions
  image_processor.py                    VaeImageProcessor.normalize / denormalize / pt_to_numpy / numpy_to_pil
PINNED (round 2): diffusers itself is not in this image and the reference holds no golden vectors for it, but the image
ships an independent executable implementation of the same network -- Black Forest Labs' FLUX `AutoEncoder`
(torchtitan/experiments/flux/model/autoencoder.py: ch 128, ch_mult (1, 2, 4, 4), z 16, scale 0.3611, shift 0.1159; the
checkpoint layout diffusers converts FLUX.1-dev's VAE from).  tests/golden/make_vae_bfl_golden.py maps this file's
diffusers-named parameters onto that module and writes tests/golden/vae_bfl_v1.npz; tests/test_vae_cpu.py replays the
fixture and re-runs the module live at the fixture's width and at FLUX.1-dev's full widths: encoder moments, sampled +
shifted + scaled latents and the decoded image agree to 1e-6 relL2.  What stays restated from memory is only the
diffusers <-> BFL key correspondence and VaeImageProcessor's normalise / denormalise / uint8 rounding.

Parameters are a flat dict in diffusers' state-dict naming ("decoder.up_blocks.0.resnets.1.conv1.weight", ...).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class VaeConfig:
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 16
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.3611
    shift_factor: float = 0.1159
    eps: float = 1e-6


def conv_shapes(cfg: VaeConfig) -> Dict[str, Tuple[int, int, int]]:
    """name -> (out_channels, in_channels, kernel) of every convolution; Linear layers of the attention use kernel 0."""
    s: Dict[str, Tuple[int, int, int]] = {}
    ch = cfg.block_out_channels

    def resnet(p, cin, cout):
        s[p + ".conv1"] = (cout, cin, 3)
        s[p + ".conv2"] = (cout, cout, 3)
        if cin != cout:
            s[p + ".conv_shortcut"] = (cout, cin, 1)

    def mid(p, c):
        resnet(p + ".resnets.0", c, c)
        resnet(p + ".resnets.1", c, c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            s[f"{p}.attentions.0.{n}"] = (c, c, 0)

    # encoder (vae.py Encoder.__init__)
    s["encoder.conv_in"] = (ch[0], cfg.in_channels, 3)
    cin = ch[0]
    for i, cout in enumerate(ch):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i != len(ch) - 1:
            s[f"encoder.down_blocks.{i}.downsamplers.0.conv"] = (cout, cout, 3)
        cin = cout
    mid("encoder.mid_block", ch[-1])
    s["encoder.conv_out"] = (2 * cfg.latent_channels, ch[-1], 3)
    # decoder (vae.py Decoder.__init__)
    s["decoder.conv_in"] = (ch[-1], cfg.latent_channels, 3)
    mid("decoder.mid_block", ch[-1])
    rev = tuple(reversed(ch))
    cin = rev[0]
    for i, cout in enumerate(rev):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
        if i != len(ch) - 1:
            s[f"decoder.up_blocks.{i}.upsamplers.0.conv"] = (cout, cout, 3)
        cin = cout
    s["decoder.conv_out"] = (cfg.out_channels, ch[0], 3)
    return s


def norm_shapes(cfg: VaeConfig) -> Dict[str, int]:
    """name -> channels of every GroupNorm."""
    n: Dict[str, int] = {}
    for name, (cout, cin, k) in conv_shapes(cfg).items():
        if name.endswith(".conv1"):
            n[name[:-5] + "norm1"] = cin
        elif name.endswith(".conv2"):
            n[name[:-5] + "norm2"] = cin
        elif name.endswith(".to_q"):
            n[name[:-4] + "group_norm"] = cin
    n["encoder.conv_norm_out"] = cfg.block_out_channels[-1]
    n["decoder.conv_norm_out"] = cfg.block_out_channels[0]
    return n


def init_params(cfg: VaeConfig, seed: int = 1234, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded synthetic parameters (no checkpoint in this image): convolutions ~ N(0, 1/fan_in) so activations keep unit
    scale through the 30-odd layers, small non-zero biases, GroupNorm weight 1 + 0.1 N(0,1), bias 0.1 N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    P: Dict[str, torch.Tensor] = {}
    for name, (cout, cin, k) in sorted(conv_shapes(cfg).items()):
        if k == 0:
            w = torch.randn(cout, cin, generator=g) / cin ** 0.5
        else:
            w = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
        P[name + ".weight"] = w.to(dtype)
        P[name + ".bias"] = (0.05 * torch.randn(cout, generator=g)).to(dtype)
    for name, c in sorted(norm_shapes(cfg).items()):
        P[name + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g)).to(dtype)
        P[name + ".bias"] = (0.1 * torch.randn(c, generator=g)).to(dtype)
    return P


# ---------------------------------------------------------------------------------------------------------------------
# blocks
# ---------------------------------------------------------------------------------------------------------------------
def _gn(P, name, x, cfg):
    return F.group_norm(x, cfg.norm_num_groups, P[name + ".weight"], P[name + ".bias"], cfg.eps)


def _conv(P, name, x, stride=1, padding=1):
    return F.conv2d(x, P[name + ".weight"], P[name + ".bias"], stride=stride, padding=padding)


def resnet_block(P, p, x, cfg):
    """ResnetBlock2D.forward with temb=None: norm1 -> silu -> conv1 -> norm2 -> silu -> conv2, + (1x1) shortcut."""
    h = _conv(P, p + ".conv1", F.silu(_gn(P, p + ".norm1", x, cfg)))
    h = _conv(P, p + ".conv2", F.silu(_gn(P, p + ".norm2", h, cfg)))
    if p + ".conv_shortcut.weight" in P:
        x = _conv(P, p + ".conv_shortcut", x, padding=0)
    return x + h


def mid_attention(P, p, x, cfg):
    """Attention.forward / AttnProcessor2_0 with one head of width C over the H*W positions, residual connection."""
    B, C, H, W = x.shape
    t = _gn(P, p + ".group_norm", x.view(B, C, H * W), cfg).transpose(1, 2)  # [B, HW, C]
    q = F.linear(t, P[p + ".to_q.weight"], P[p + ".to_q.bias"])
    k = F.linear(t, P[p + ".to_k.weight"], P[p + ".to_k.bias"])
    v = F.linear(t, P[p + ".to_v.weight"], P[p + ".to_v.bias"])
    a = torch.softmax(q @ k.transpose(1, 2) * (C ** -0.5), dim=-1) @ v
    o = F.linear(a, P[p + ".to_out.0.weight"], P[p + ".to_out.0.bias"])
    return o.transpose(1, 2).reshape(B, C, H, W) + x


def mid_block(P, p, x, cfg):
    x = resnet_block(P, p + ".resnets.0", x, cfg)
    x = mid_attention(P, p + ".attentions.0", x, cfg)
    return resnet_block(P, p + ".resnets.1", x, cfg)


def encode_moments(P: Dict[str, torch.Tensor], images: torch.Tensor, cfg: VaeConfig) -> torch.Tensor:
    """Encoder.forward: images [B, 3, H, W] in [-1, 1] -> moments [B, 2*latent, H/8, W/8] (mean | logvar)."""
    x = _conv(P, "encoder.conv_in", images)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block):
            x = resnet_block(P, f"encoder.down_blocks.{i}.resnets.{j}", x, cfg)
        if i != n - 1:
            x = _conv(P, f"encoder.down_blocks.{i}.downsamplers.0.conv", F.pad(x, (0, 1, 0, 1)), stride=2, padding=0)
    x = mid_block(P, "encoder.mid_block", x, cfg)
    return _conv(P, "encoder.conv_out", F.silu(_gn(P, "encoder.conv_norm_out", x, cfg)))


def sample_latents(moments: torch.Tensor, eps: Optional[torch.Tensor]) -> torch.Tensor:
    """DiagonalGaussianDistribution: mean + exp(0.5 * clamp(logvar, -30, 20)) * eps   (eps None -> the mode)."""
    mean, logvar = moments.chunk(2, dim=1)
    if eps is None:
        return mean
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * eps


def encode(P, images, cfg, eps=None) -> torch.Tensor:
    """pipeline_tools.py:9-12: sample, then (z - shift) * scale."""
    z = sample_latents(encode_moments(P, images, cfg), eps)
    return (z - cfg.shift_factor) * cfg.scaling_factor


def decode_raw(P: Dict[str, torch.Tensor], z: torch.Tensor, cfg: VaeConfig) -> torch.Tensor:
    """Decoder.forward: z [B, latent, h, w] -> image [B, 3, 8h, 8w]."""
    x = _conv(P, "decoder.conv_in", z)
    x = mid_block(P, "decoder.mid_block", x, cfg)
    n = len(cfg.block_out_channels)
    for i in range(n):
        for j in range(cfg.layers_per_block + 1):
            x = resnet_block(P, f"decoder.up_blocks.{i}.resnets.{j}", x, cfg)
        if i != n - 1:
            x = _conv(P, f"decoder.up_blocks.{i}.upsamplers.0.conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))
    return _conv(P, "decoder.conv_out", F.silu(_gn(P, "decoder.conv_norm_out", x, cfg)))


def decode(P, latents, cfg) -> torch.Tensor:
    """generate.py:376-379: z / scale + shift, then the decoder."""
    return decode_raw(P, latents / cfg.scaling_factor + cfg.shift_factor, cfg)


# ---------------------------------------------------------------------------------------------------------------------
# VaeImageProcessor (image_processor.py)
# ---------------------------------------------------------------------------------------------------------------------
def preprocess(images01: torch.Tensor) -> torch.Tensor:
    """normalize: [0, 1] -> [-1, 1]."""
    return 2.0 * images01 - 1.0


def postprocess_pt(image: torch.Tensor) -> torch.Tensor:
    """denormalize: (x / 2 + 0.5).clamp(0, 1)."""
    return (image / 2 + 0.5).clamp(0, 1)


def postprocess_uint8(image: torch.Tensor) -> torch.Tensor:
    """pt_to_numpy + numpy_to_pil's quantisation: [B, H, W, 3] uint8 = round(255 * denormalized)."""
    return (postprocess_pt(image).permute(0, 2, 3, 1).float() * 255).round().to(torch.uint8)
