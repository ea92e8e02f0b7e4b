"""Native FLUX VAE: `pipeline.vae` / `pipeline.image_processor` for the two call sites either side of
the denoising loop — pipeline_tools.py:7-12 (`vae.encode(images).latent_dist.sample()`) and generate.py:375-380
(`vae.decode(z, return_dict=False)[0]`, `image_processor.postprocess`).

The arithmetic is diffusers 0.31.0's AutoencoderKL with FLUX.1-dev's vae/config.json (restated in oracle/vae.py, which
this module never imports).  Host side = pointer plumbing: every convolution is `lx_vae_im2col` (GroupNorm + SiLU +
up-sampling / stride folded into the panel write) + the tcgen05 GEMM; the mid-block attention (one head of width 512
over H*W positions) is three GEMMs around a row softmax with fp32 logits.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch

from . import _lib as L
from . import ops

_lib = L.lib
c_void_p, c_int32, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class Im2colDesc(C.Structure):
    _fields_ = [("x", c_void_p), ("out", c_void_p), ("coeff", c_void_p), ("B", c_int32), ("H", c_int32), ("W", c_int32),
                ("C", c_int32), ("upsample", c_int32), ("stride", c_int32), ("pad_lo", c_int32), ("taps", c_int32),
                ("silu", c_int32), ("Ho", c_int32), ("Wo", c_int32), ("ldk", c_int64)]


_lib.lx_vae_group_norm_coeffs.argtypes = [c_void_p, c_int32, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_float, c_void_p,
                                          c_void_p, c_void_p]
_lib.lx_vae_group_norm_workspace.argtypes = [c_int32, c_int64, c_int32]
_lib.lx_vae_group_norm_workspace.restype = c_int64
_lib.lx_vae_im2col.argtypes = [C.POINTER(Im2colDesc), c_void_p]
_lib.lx_vae_softmax_rows.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_float, c_void_p]
_lib.lx_vae_nchw_to_rows.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int32, c_float, c_float, c_void_p]
_lib.lx_vae_rows_to_nchw.argtypes = [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_int64, c_int32, c_void_p]
_lib.lx_vae_sample_latents.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int64, c_float, c_float,
                                       c_void_p]
_lib.lx_transpose_bf16.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p]


def _stream() -> int:
    return L.current_stream()


def _cuda(t: torch.Tensor) -> int:
    assert t.is_cuda, "loongx_b200.vae needs CUDA tensors (there is no CPU fallback)"
    return t.data_ptr()


def _up(n: int, m: int) -> int:
    return (n + m - 1) // m * m


@dataclass
class VaeConfig:
    """FLUX.1-dev vae/config.json (diffusers AutoencoderKL)."""
    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 16
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.3611
    shift_factor: float = 0.1159
    eps: float = 1e-6

    @staticmethod
    def from_json(cj: dict) -> "VaeConfig":
        kw = {k: cj[k] for k in ("in_channels", "out_channels", "latent_channels", "layers_per_block", "norm_num_groups",
                                 "scaling_factor", "shift_factor") if k in cj and cj[k] is not None}
        if "block_out_channels" in cj:
            kw["block_out_channels"] = tuple(cj["block_out_channels"])
        if cj.get("use_quant_conv") or cj.get("use_post_quant_conv"):
            raise NotImplementedError("quant_conv / post_quant_conv are not part of the FLUX VAE and not built")
        if cj.get("mid_block_add_attention") is False:
            raise NotImplementedError("mid_block_add_attention=False is not built")
        return VaeConfig(**kw)


# ------------------------------------------------------------------------------------------------------------------
# layer tables (integer bookkeeping only; mirrors AutoencoderKL's module tree)
# ------------------------------------------------------------------------------------------------------------------
def conv_table(cfg: VaeConfig) -> Dict[str, Tuple[int, int, int]]:
    """diffusers module name -> (out_channels, in_channels, kernel); kernel 0 marks the attention's Linear layers."""
    t: Dict[str, Tuple[int, int, int]] = {}
    ch = tuple(cfg.block_out_channels)

    def resnet(p, cin, cout):
        t[p + ".conv1"] = (cout, cin, 3)
        t[p + ".conv2"] = (cout, cout, 3)
        if cin != cout:
            t[p + ".conv_shortcut"] = (cout, cin, 1)

    def mid(p, c):
        for j in (0, 1):
            resnet(f"{p}.resnets.{j}", c, c)
        for n in ("to_q", "to_k", "to_v", "to_out.0"):
            t[f"{p}.attentions.0.{n}"] = (c, c, 0)

    t["encoder.conv_in"] = (ch[0], cfg.in_channels, 3)
    prev = ch[0]
    for i, c in enumerate(ch):
        for j in range(cfg.layers_per_block):
            resnet(f"encoder.down_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        if i + 1 < len(ch):
            t[f"encoder.down_blocks.{i}.downsamplers.0.conv"] = (c, c, 3)
        prev = c
    mid("encoder.mid_block", ch[-1])
    t["encoder.conv_out"] = (2 * cfg.latent_channels, ch[-1], 3)
    t["decoder.conv_in"] = (ch[-1], cfg.latent_channels, 3)
    mid("decoder.mid_block", ch[-1])
    prev = ch[-1]
    for i, c in enumerate(reversed(ch)):
        for j in range(cfg.layers_per_block + 1):
            resnet(f"decoder.up_blocks.{i}.resnets.{j}", prev if j == 0 else c, c)
        if i + 1 < len(ch):
            t[f"decoder.up_blocks.{i}.upsamplers.0.conv"] = (c, c, 3)
        prev = c
    t["decoder.conv_out"] = (cfg.out_channels, ch[0], 3)
    return t


def norm_table(cfg: VaeConfig) -> Dict[str, int]:
    n: Dict[str, int] = {}
    for name, (cout, cin, k) in conv_table(cfg).items():
        stem, leaf = name.rsplit(".", 1)
        if leaf == "conv1":
            n[stem + ".norm1"] = cin
        elif leaf == "conv2":
            n[stem + ".norm2"] = cin
        elif leaf == "to_q":
            n[stem + ".group_norm"] = cin
    n["encoder.conv_norm_out"] = cfg.block_out_channels[-1]
    n["decoder.conv_norm_out"] = cfg.block_out_channels[0]
    return n


def expected_keys(cfg: VaeConfig) -> set:
    keys = set()
    for name in list(conv_table(cfg)) + list(norm_table(cfg)):
        keys.add(name + ".weight")
        keys.add(name + ".bias")
    return keys


def pack_conv(w: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, int, int]:
    """Conv2d weight [Cout, Cin, k, k] (or Linear [Cout, Cin]) -> (bf16 panel [Cout_pad, ldk] with columns ordered
    (ky, kx, c) over the channel count padded to 8, fp32 bias [Cout_pad], taps, padded Cin).  ldk is a multiple of 64
    (one GEMM K block); the padding is zeros on both operands."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    cout, cin, kh, kw = w.shape
    assert kh == kw and kh in (1, 3)
    cin_p, cout_p = _up(cin, 8), _up(cout, 8)
    taps = kh * kw
    ldk = _up(taps * cin_p, 64)
    panel = torch.zeros(cout_p, ldk, dtype=torch.float32, device=w.device)
    wp = torch.zeros(cout, kh, kw, cin_p, dtype=torch.float32, device=w.device)
    wp[..., :cin] = w.float().permute(0, 2, 3, 1)
    panel[:cout, :taps * cin_p] = wp.reshape(cout, taps * cin_p)
    bias = torch.zeros(cout_p, dtype=torch.float32, device=w.device)
    bias[:cout] = b.float()
    return panel.to(torch.bfloat16).contiguous(), bias, taps, cin_p


@dataclass
class _Conv:
    w: torch.Tensor
    bias: torch.Tensor
    taps: int
    cin: int   # padded to 8
    cout: int  # padded to 8


class VaeWeights:
    """Packed parameters of one AutoencoderKL on the device (diffusers state-dict naming in, GEMM panels out)."""

    def __init__(self, cfg: VaeConfig, P: Dict[str, torch.Tensor], device="cuda"):
        missing = sorted(expected_keys(cfg) - set(P))
        extra = sorted(set(P) - expected_keys(cfg))
        if missing or extra:
            raise KeyError(f"VAE checkpoint does not match the config: missing {missing[:4]} ({len(missing)}), "
                           f"unexpected {extra[:4]} ({len(extra)})")
        self.cfg = cfg
        self.device = torch.device(device)
        self.conv: Dict[str, _Conv] = {}
        self.norm: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}
        table = conv_table(cfg)
        for name, (cout, cin, k) in table.items():
            w = P[name + ".weight"].to(self.device)
            if tuple(w.shape[:2]) != (cout, cin):
                raise ValueError(f"{name}: expected [{cout}, {cin}, ...], got {tuple(w.shape)}")
            if name.endswith((".to_k", ".to_v")):
                continue  # folded into the to_q entry below
            if name.endswith(".to_q"):
                stem = name[:-5]
                w = torch.cat([P[f"{stem}.to_{n}.weight"].to(self.device) for n in "qkv"], 0)
                b = torch.cat([P[f"{stem}.to_{n}.bias"].to(self.device) for n in "qkv"], 0)
                name = stem + ".to_qkv"
            else:
                b = P[name + ".bias"].to(self.device)
            panel, bias, taps, cin_p = pack_conv(w, b)
            self.conv[name] = _Conv(panel, bias, taps, cin_p, panel.shape[0])
        for name in norm_table(cfg):
            self.norm[name] = (P[name + ".weight"].to(self.device, torch.float32).contiguous(),
                               P[name + ".bias"].to(self.device, torch.float32).contiguous())

    @staticmethod
    def from_pretrained(flux_path: str, device="cuda") -> "VaeWeights":
        """<flux_path>/vae/{config.json, diffusion_pytorch_model.safetensors} (FluxPipeline.from_pretrained layout)."""
        from safetensors import safe_open

        vdir = os.path.join(flux_path, "vae") if os.path.isdir(os.path.join(flux_path, "vae")) else flux_path
        with open(os.path.join(vdir, "config.json")) as f:
            cfg = VaeConfig.from_json(json.load(f))
        P = {}
        with safe_open(os.path.join(vdir, "diffusion_pytorch_model.safetensors"), framework="pt", device="cpu") as sf:
            for k in sf.keys():
                P[k] = sf.get_tensor(k)
        return VaeWeights(cfg, P, device)


def write_diffusers_vae(path: str, cfg: VaeConfig, P: Dict[str, torch.Tensor]) -> None:
    """Write <path>/vae/{config.json, diffusion_pytorch_model.safetensors} in the layout FluxPipeline.from_pretrained reads
    (used by the tests and to export synthetic weights)."""
    from safetensors.torch import save_file

    vdir = os.path.join(path, "vae")
    os.makedirs(vdir, exist_ok=True)
    cj = {"_class_name": "AutoencoderKL", "in_channels": cfg.in_channels, "out_channels": cfg.out_channels,
          "latent_channels": cfg.latent_channels, "block_out_channels": list(cfg.block_out_channels),
          "layers_per_block": cfg.layers_per_block, "norm_num_groups": cfg.norm_num_groups, "act_fn": "silu",
          "scaling_factor": cfg.scaling_factor, "shift_factor": cfg.shift_factor, "use_quant_conv": False,
          "use_post_quant_conv": False, "mid_block_add_attention": True, "force_upcast": True}
    with open(os.path.join(vdir, "config.json"), "w") as f:
        json.dump(cj, f, indent=1)
    save_file({k: v.contiguous().cpu() for k, v in P.items()}, os.path.join(vdir, "diffusion_pytorch_model.safetensors"))


def synthetic_params(cfg: VaeConfig, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Seeded random parameters of the right shapes (no checkpoint in this image).  Same recipe as the oracle's
    init_params so the tests can build both sides from one seed: convolutions ~ N(0, 1/fan_in), biases 0.05 N(0,1),
    GroupNorm weight 1 + 0.1 N(0,1), bias 0.1 N(0,1), drawn in sorted-name order on the CPU generator."""
    g = torch.Generator().manual_seed(seed)
    P: Dict[str, torch.Tensor] = {}
    for name, (cout, cin, k) in sorted(conv_table(cfg).items()):
        shape = (cout, cin) if k == 0 else (cout, cin, k, k)
        fan_in = cin * max(k, 1) ** 2
        P[name + ".weight"] = torch.randn(*shape, generator=g) / fan_in ** 0.5
        P[name + ".bias"] = 0.05 * torch.randn(cout, generator=g)
    for name, c in sorted(norm_table(cfg).items()):
        P[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        P[name + ".bias"] = 0.1 * torch.randn(c, generator=g)
    return P


# ------------------------------------------------------------------------------------------------------------------
# execution
# ------------------------------------------------------------------------------------------------------------------
@dataclass
class _Act:
    """bf16 rows [B*H*W, C] of an NHWC activation."""
    x: torch.Tensor
    B: int
    H: int
    W: int

    @property
    def C(self) -> int:
        return self.x.shape[1]


class _LatentDist:
    """The slice of DiagonalGaussianDistribution the reference touches (`.sample()`, pipeline_tools.py:9), plus `.mode()`."""

    def __init__(self, moments: torch.Tensor, B: int, h: int, w: int, L: int, dtype):
        self._m, self._B, self._h, self._w, self._L, self._dtype = moments, B, h, w, L, dtype

    def _draw(self, eps: Optional[torch.Tensor]) -> torch.Tensor:
        out = torch.empty(self._B, self._L, self._h, self._w, dtype=torch.float32, device=self._m.device)
        L.check(_lib.lx_vae_sample_latents(_cuda(self._m), self._m.stride(0), None if eps is None else _cuda(eps), _cuda(out),
                                           self._B, self._L, self._h * self._w, 0.0, 1.0, _stream()), "lx_vae_sample_latents")
        return out.to(self._dtype)

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        eps = torch.randn(self._B, self._L, self._h, self._w, generator=generator, device=self._m.device, dtype=torch.float32)
        return self._draw(eps)

    def mode(self) -> torch.Tensor:
        return self._draw(None)


@dataclass
class _EncoderOutput:
    latent_dist: _LatentDist


@dataclass
class _DecoderOutput:
    sample: torch.Tensor


@dataclass
class _VaeConfigView:
    scaling_factor: float
    shift_factor: float
    latent_channels: int
    block_out_channels: Tuple[int, ...]
    extra: dict = field(default_factory=dict)


class NativeVae:
    """`pipeline.vae`: `.config.{scaling_factor, shift_factor}`, `.encode(images).latent_dist.sample()`, `.decode(z)`."""

    PANEL_BYTES = 6 << 30  # im2col panels of one pass are kept under this by splitting the batch

    def __init__(self, weights: VaeWeights):
        self.w = weights
        cfg = weights.cfg
        self.cfg = cfg
        self.device = weights.device
        self.dtype = torch.bfloat16
        self.config = _VaeConfigView(cfg.scaling_factor, cfg.shift_factor, cfg.latent_channels, tuple(cfg.block_out_channels))
        self._ones = torch.ones(4096, dtype=torch.bfloat16, device=self.device)
        self._meta = torch.zeros(0, 4, dtype=torch.int32, device=self.device)
        self.launches = 0

    # -- primitives ------------------------------------------------------------------------------------------------
    def _tile_meta(self, M: int) -> torch.Tensor:
        need = (M + 127) // 128 + 2
        if self._meta.shape[0] < need:  # all zeros: every tile is (stream 0, batch 0) -> the gate is the ones vector
            self._meta = torch.zeros(need, 4, dtype=torch.int32, device=self.device)
        return self._meta

    def _coeffs(self, a: _Act, norm: str) -> torch.Tensor:
        gamma, beta = self.w.norm[norm]
        assert gamma.numel() == a.C, f"{norm}: {gamma.numel()} channels, activation has {a.C}"
        g = self.cfg.norm_num_groups
        sums = torch.empty(_lib.lx_vae_group_norm_workspace(a.B, a.H * a.W, g), dtype=torch.float64, device=self.device)
        coeff = torch.empty(a.B, a.C, 2, dtype=torch.float32, device=self.device)
        L.check(_lib.lx_vae_group_norm_coeffs(_cuda(a.x), a.B, a.H * a.W, a.C, g, _cuda(gamma), _cuda(beta), self.cfg.eps,
                                              _cuda(sums), _cuda(coeff), _stream()), "lx_vae_group_norm_coeffs")
        self.launches += 2
        return coeff

    def _panel(self, a: _Act, ldk: int, taps: int, coeff=None, silu=False, up=1, stride=1, pad_lo=1):
        if taps == 1:
            Ho, Wo, pad_lo = a.H, a.W, 0
        elif stride == 1:
            Ho, Wo = a.H * up, a.W * up
        else:  # Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then a 3x3 stride-2 convolution
            Ho, Wo = (a.H + 1 - 3) // 2 + 1, (a.W + 1 - 3) // 2 + 1
        out = torch.empty(a.B * Ho * Wo, ldk, dtype=torch.bfloat16, device=self.device)
        d = Im2colDesc()
        d.x, d.out, d.coeff = _cuda(a.x), _cuda(out), None if coeff is None else _cuda(coeff)
        d.B, d.H, d.W, d.C = a.B, a.H, a.W, a.C
        d.upsample, d.stride, d.pad_lo, d.taps, d.silu = up, stride, pad_lo, taps, int(bool(silu))
        d.Ho, d.Wo, d.ldk = Ho, Wo, ldk
        L.check(_lib.lx_vae_im2col(C.byref(d), _stream()), "lx_vae_im2col")
        self.launches += 1
        return out, Ho, Wo

    def _gemm(self, A, conv: _Conv, residual: Optional[torch.Tensor] = None, f32: bool = False) -> torch.Tensor:
        M = A.shape[0]
        out = torch.empty(M, conv.cout, dtype=torch.float32 if f32 else torch.bfloat16, device=self.device)
        tile_n = 128 if conv.cout <= 128 else 0  # the 128-channel layers hold half of the decoder's FLOPs
        if residual is not None:
            assert not f32 and residual.shape == out.shape
            ops.gemm(A, conv.w, conv.bias, out, L.EPI_GATE_RESIDUAL, tile_meta=self._tile_meta(M), residual=residual,
                     gate=[self._ones[:conv.cout], None, None], tile_n=tile_n)
        else:
            ops.gemm(A, conv.w, conv.bias, out, L.EPI_BIAS_F32 if f32 else L.EPI_BIAS, tile_n=tile_n)
        self.launches += 1
        return out

    def _conv(self, name: str, a: _Act, norm: Optional[str] = None, silu=False, up=1, stride=1, pad_lo=1,
              residual: Optional[torch.Tensor] = None, f32=False) -> _Act:
        cv = self.w.conv[name]
        assert cv.cin == a.C, f"{name}: expects {cv.cin} input channels, activation has {a.C}"
        coeff = self._coeffs(a, norm) if norm is not None else None
        if cv.taps == 1 and coeff is None:
            A, Ho, Wo = a.x, a.H, a.W  # 1x1 convolution of the stored activation: the rows are the GEMM operand
            assert cv.w.shape[1] == a.C
        else:
            A, Ho, Wo = self._panel(a, cv.w.shape[1], cv.taps, coeff, silu, up, stride, pad_lo)
        return _Act(self._gemm(A, cv, residual, f32), a.B, Ho, Wo)

    def _resnet(self, p: str, a: _Act) -> _Act:
        """ResnetBlock2D (temb None): conv2(silu(norm2(conv1(silu(norm1(x)))))) + shortcut(x)."""
        h = self._conv(p + ".conv1", a, norm=p + ".norm1", silu=True)
        res = self._conv(p + ".conv_shortcut", a).x if (p + ".conv_shortcut") in self.w.conv else a.x
        return self._conv(p + ".conv2", h, norm=p + ".norm2", silu=True, residual=res)

    def _attention(self, p: str, a: _Act) -> _Act:
        """Attention(heads=1, dim_head=C, residual_connection=True) over the H*W positions, fp32 logits."""
        B, hw, Cc = a.B, a.H * a.W, a.C
        n_pad = _up(hw, 8)
        qkv_w, out_w = self.w.conv[p + ".to_qkv"], self.w.conv[p + ".to_out.0"]
        t, _, _ = self._panel(a, Cc, 1, self._coeffs(a, p + ".group_norm"), silu=False)
        qkv = torch.zeros(B * hw + 8, 3 * Cc, dtype=torch.bfloat16, device=self.device)
        ops.gemm(t, qkv_w.w, qkv_w.bias, qkv[:B * hw], L.EPI_BIAS)
        attn = torch.empty(B * hw, Cc, dtype=torch.bfloat16, device=self.device)
        S = torch.empty(hw, n_pad, dtype=torch.float32, device=self.device)
        Pm = torch.empty(hw, n_pad, dtype=torch.bfloat16, device=self.device)
        vT = torch.zeros(Cc, n_pad, dtype=torch.bfloat16, device=self.device)
        for b in range(B):
            r0 = b * hw
            q, k, v = qkv[r0:r0 + hw, :Cc], qkv[r0:r0 + n_pad, Cc:2 * Cc], qkv[r0:r0 + hw, 2 * Cc:]
            ops.gemm(q, k, None, S, L.EPI_BIAS_F32, w_dynamic=True)  # W = this sample's keys
            L.check(_lib.lx_vae_softmax_rows(_cuda(S), S.stride(0), _cuda(Pm), Pm.stride(0), hw, hw, float(Cc) ** -0.5,
                                             _stream()), "lx_vae_softmax_rows")
            L.check(_lib.lx_transpose_bf16(_cuda(v), v.stride(0), _cuda(vT), vT.stride(0), hw, Cc, _stream()),
                    "lx_transpose_bf16")
            ops.gemm(Pm, vT, None, attn[r0:r0 + hw], L.EPI_BIAS, w_dynamic=True)  # W = V^T
        self.launches += 1 + 4 * B
        return _Act(self._gemm(attn, out_w, residual=a.x), B, a.H, a.W)

    def _mid(self, p: str, a: _Act) -> _Act:
        a = self._resnet(p + ".resnets.0", a)
        a = self._attention(p + ".attentions.0", a)
        return self._resnet(p + ".resnets.1", a)

    def _rows_in(self, x: torch.Tensor, c_pad: int, mul: float, add: float) -> _Act:
        B, Cc, H, W = x.shape
        x = x.to(self.device, torch.float32).contiguous()
        rows = torch.empty(B * H * W, c_pad, dtype=torch.bfloat16, device=self.device)
        L.check(_lib.lx_vae_nchw_to_rows(_cuda(x), _cuda(rows), B, Cc, H * W, c_pad, mul, add, _stream()), "lx_vae_nchw_to_rows")
        self.launches += 1
        return _Act(rows, B, H, W)

    def _split(self, B: int, out_pixels: int, widest: int) -> int:
        """samples per pass so that the largest im2col panel stays under PANEL_BYTES."""
        per_sample = out_pixels * 9 * widest * 2
        return max(1, min(B, self.PANEL_BYTES // max(per_sample, 1)))

    # -- AutoencoderKL surface ---------------------------------------------------------------------------------------
    def decode_rows(self, z: torch.Tensor) -> Tuple[torch.Tensor, int, int, int]:
        """z [B, latent, h, w] -> fp32 rows [B*8h*8w, 8] (channels 0..2 are the image), B, H, W."""
        cfg = self.cfg
        a = self._rows_in(z, self.w.conv["decoder.conv_in"].cin, 1.0, 0.0)
        a = self._conv("decoder.conv_in", a)
        a = self._mid("decoder.mid_block", a)
        n = len(cfg.block_out_channels)
        for i in range(n):
            for j in range(cfg.layers_per_block + 1):
                a = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", a)
            if i + 1 < n:
                a = self._conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", a, up=2)
        a = self._conv("decoder.conv_out", a, norm="decoder.conv_norm_out", silu=True, f32=True)
        return a.x, a.B, a.H, a.W

    def decode(self, z: torch.Tensor, return_dict: bool = True, generator=None):
        assert z.dim() == 4 and z.shape[1] == self.cfg.latent_channels, f"latents must be [B, {self.cfg.latent_channels}, h, w]"
        B, _, h, w = z.shape
        up = 2 ** (len(self.cfg.block_out_channels) - 1)
        out = torch.empty(B, self.cfg.out_channels, h * up, w * up, dtype=torch.float32, device=self.device)
        step = self._split(B, h * up * w * up, self.cfg.block_out_channels[1])
        for b0 in range(0, B, step):
            rows, nb, H, W = self.decode_rows(z[b0:b0 + step])
            L.check(_lib.lx_vae_rows_to_nchw(_cuda(rows), rows.stride(0), _cuda(out[b0:b0 + nb]), nb, self.cfg.out_channels,
                                             H * W, 0, _stream()), "lx_vae_rows_to_nchw")
            self.launches += 1
        image = out.to(z.dtype) if z.dtype in (torch.bfloat16, torch.float16) else out
        return _DecoderOutput(image) if return_dict else (image,)

    def encode_moments(self, images: torch.Tensor) -> Tuple[torch.Tensor, int, int, int]:
        """images [B, 3, H, W] in [-1, 1] -> fp32 moment rows [B*h*w, 2*latent] (mean | logvar), B, h, w."""
        cfg = self.cfg
        a = self._rows_in(images, self.w.conv["encoder.conv_in"].cin, 1.0, 0.0)
        a = self._conv("encoder.conv_in", a)
        n = len(cfg.block_out_channels)
        for i in range(n):
            for j in range(cfg.layers_per_block):
                a = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", a)
            if i + 1 < n:
                a = self._conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", a, stride=2, pad_lo=0)
        a = self._mid("encoder.mid_block", a)
        a = self._conv("encoder.conv_out", a, norm="encoder.conv_norm_out", silu=True, f32=True)
        return a.x, a.B, a.H, a.W

    def encode(self, images: torch.Tensor, return_dict: bool = True):
        assert images.dim() == 4 and images.shape[1] == self.cfg.in_channels, "images must be [B, 3, H, W]"
        f = 2 ** (len(self.cfg.block_out_channels) - 1)
        B, _, H, W = images.shape
        if H % f or W % f:
            raise ValueError(f"image size {H}x{W} must be a multiple of {f}")
        step = self._split(B, H * W, self.cfg.block_out_channels[0])
        parts = [self.encode_moments(images[b0:b0 + step])[0] for b0 in range(0, B, step)]
        moments = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
        dist = _LatentDist(moments, B, H // f, W // f, self.cfg.latent_channels, images.dtype)
        return _EncoderOutput(dist) if return_dict else (dist,)


class ImageProcessor:
    """The slice of diffusers' VaeImageProcessor the reference touches: `preprocess` (pipeline_tools.py:8) and
    `postprocess(image, output_type)` (generate.py:380)."""

    def __init__(self, vae_scale_factor: int = 16):
        self.vae_scale_factor = vae_scale_factor

    def preprocess(self, images) -> torch.Tensor:
        """PIL image(s) / [0, 1] tensors -> fp32 [B, 3, H, W] in [-1, 1] (VaeImageProcessor.normalize).  A tensor that
        already holds negative values is passed through, as diffusers does."""
        if isinstance(images, torch.Tensor):
            x = images if images.dim() == 4 else images[None]
            f = self.vae_scale_factor
            if x.shape[-2] % f or x.shape[-1] % f:  # VaeImageProcessor resizes tensors too; here: be explicit
                raise ValueError(f"tensor images must have H, W divisible by {f} (got {tuple(x.shape[-2:])}); resize them "
                                 "first (PIL inputs are resized like VaeImageProcessor does)")
            return x if x.min() < 0 else 2.0 * x - 1.0
        import numpy as np

        if not isinstance(images, (list, tuple)):
            images = [images]
        f = self.vae_scale_factor
        fitted = []
        for im in images:  # VaeImageProcessor(do_resize=True): down to the next multiple of vae_scale_factor, lanczos
            w, h = im.size
            w2, h2 = w - w % f, h - h % f
            if w2 == 0 or h2 == 0:
                raise ValueError(f"image size {w}x{h} is smaller than the VAE scale factor {f}")
            if (w2, h2) != (w, h):
                from PIL import Image

                im = im.resize((w2, h2), resample=Image.LANCZOS)
            fitted.append(np.asarray(im.convert("RGB"), dtype=np.float32) / 255.0)
        if len({a.shape for a in fitted}) != 1:
            raise ValueError("images of one batch must have the same size")
        arr = np.stack(fitted, 0)
        return 2.0 * torch.from_numpy(arr).permute(0, 3, 1, 2).contiguous() - 1.0

    def postprocess(self, image: torch.Tensor, output_type: str = "pil"):
        if output_type == "latent":
            return image
        x = image.float().contiguous()
        out = torch.empty_like(x)
        # (x / 2 + 0.5).clamp(0, 1): the denormalising copy kernel over the flat tensor
        L.check(_lib.lx_vae_rows_to_nchw(_cuda(x), 1, _cuda(out), 1, 1, x.numel(), 1, _stream()), "lx_vae_rows_to_nchw")
        if output_type == "pt":
            return out
        arr = out.permute(0, 2, 3, 1).cpu().numpy()
        if output_type == "np":
            return arr
        if output_type == "pil":
            from PIL import Image

            return [Image.fromarray((a * 255).round().astype("uint8")) for a in arr]
        raise ValueError(f"unknown output_type {output_type!r}")
