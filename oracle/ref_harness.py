"""ORACLE — test infrastructure only (see oracle/flux_dit.py header for who may import this).

Runs the reference's OWN Python source, imported from /root/reference where it lies, on the CPU — with the third-party
packages it imports (diffusers 0.31.0, peft, accelerate, lightning, prodigyopt, s4torch: none installed here, no
network) replaced by minimal stand-ins.  What executes for real:

    /root/reference/src/flux/block.py           attn_forward, block_forward, single_block_forward
    /root/reference/src/flux/transformer.py     prepare_params, tranformer_forward
    /root/reference/src/flux/lora_controller.py enable_lora, set_lora_scale
    /root/reference/src/flux/generate.py        generate (prepare_params, conditioning glue, denoise loop)
    /root/reference/src/flux/condition.py       Condition.encode (id arithmetic)
    /root/reference/src/flux/pipeline_tools.py  encode_images (id fallback)
    /root/reference/src/train/model.py          EEGEncoder, PPGEncoder, FNIRSEncoder, MotionEncoder,
                                                FeaturePyramidPooling, DUAN, OminiModel.spatial_pyramid_pooling,
                                                OminiModel.fuse_eeg / fuse_fnirs

What is a stand-in (a second, module-shaped restatement of SURVEY.md App. A / B, independent of the functional one in
oracle/flux_dit.py, so the two cross-check each other): the diffusers classes the reference *receives* as arguments
(Attention, AdaLayerNormZero/-Single/-Continuous, RMSNorm, FeedForward, FluxPosEmbed, the timestep/text embedders,
FlowMatchEulerDiscreteScheduler, FluxPipeline's pack / ids / prepare_latents helpers), peft's LoRA Linear
(BaseTunerLayer surface: scaling / active_adapters / scale_layer) and s4torch.S4Model (= oracle.cs3_dgf.S4Model).

This module can only run where /root/reference exists (this container, not the GPU box); tests/golden/make_ref_golden.py
uses it to write tests/golden/ref_v1.npz, and tests/test_reference_pins_cpu.py re-runs it live when the tree is there.
"""
from __future__ import annotations

import contextlib
import importlib
import logging
import math
import os
import sys
import types
from typing import Dict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = "/root/reference"
PKG = "loongx_reference"  # alias package name for /root/reference/src (our own repo already owns `src`)


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "flux", "block.py"))


# ------------------------------------------------------------------------------------------------------------------
# peft stand-in
# ------------------------------------------------------------------------------------------------------------------
class BaseTunerLayer:
    """Surface used by lora_controller.py:10-42: isinstance check, .scaling, .active_adapters, .scale_layer."""

    active_adapters = ("default",)

    def scale_layer(self, scale: float) -> None:
        if scale == 1:
            return
        for a in self.active_adapters:
            self.scaling[a] *= scale


class LoraLinear(nn.Module, BaseTunerLayer):
    """peft.tuners.lora.Linear: y = base(x) + lora_B(lora_A(dropout(x))) * scaling, scaling = lora_alpha / r."""

    def __init__(self, base: nn.Linear, r: int, alpha: float):
        super().__init__()
        self.base_layer = base
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False)})
        self.scaling = {"default": alpha / r}

    def forward(self, x):
        y = self.base_layer(x)
        for a in self.active_adapters:
            y = y + self.lora_B[a](self.lora_A[a](x)) * self.scaling[a]
        return y


# ------------------------------------------------------------------------------------------------------------------
# diffusers 0.31.0 stand-ins
# ------------------------------------------------------------------------------------------------------------------
class RMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        in_dtype = x.dtype
        var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(var + self.eps)
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            x = x.to(self.weight.dtype)
        x = x * self.weight
        return x if self.weight.dtype in (torch.float16, torch.bfloat16) else x.to(in_dtype)


class AdaLayerNormZero(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(D, 6 * D)
        self.norm = nn.LayerNorm(D, elementwise_affine=False, eps=1e-6)

    def forward(self, x, timestep=None, class_labels=None, hidden_dtype=None, emb=None):
        emb = self.linear(self.silu(emb))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
        x = self.norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
        return x, gate_msa, shift_mlp, scale_mlp, gate_mlp


class AdaLayerNormZeroSingle(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(D, 3 * D)
        self.norm = nn.LayerNorm(D, elementwise_affine=False, eps=1e-6)

    def forward(self, x, emb=None):
        emb = self.linear(self.silu(emb))
        shift_msa, scale_msa, gate_msa = emb.chunk(3, dim=1)
        x = self.norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
        return x, gate_msa


class AdaLayerNormContinuous(nn.Module):
    def __init__(self, D):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(D, 2 * D)
        self.norm = nn.LayerNorm(D, elementwise_affine=False, eps=1e-6)

    def forward(self, x, conditioning_embedding):
        emb = self.linear(self.silu(conditioning_embedding).to(x.dtype))
        scale, shift = torch.chunk(emb, 2, dim=1)
        return self.norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]


class Attention(nn.Module):
    """Only the attribute surface block.py:7-176 touches (the processor is replaced by attn_forward)."""

    def __init__(self, D, heads, head_dim, pre_only=False):
        super().__init__()
        self.heads = heads
        self.to_q, self.to_k, self.to_v = nn.Linear(D, D), nn.Linear(D, D), nn.Linear(D, D)
        self.norm_q, self.norm_k = RMSNorm(head_dim), RMSNorm(head_dim)
        if not pre_only:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = nn.Linear(D, D), nn.Linear(D, D), nn.Linear(D, D)
            self.norm_added_q, self.norm_added_k = RMSNorm(head_dim), RMSNorm(head_dim)
            self.to_out = nn.ModuleList([nn.Linear(D, D), nn.Dropout(0.0)])
            self.to_add_out = nn.Linear(D, D)


class GELU(nn.Module):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.proj = nn.Linear(d_in, d_out)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForward(nn.Module):
    def __init__(self, D, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GELU(D, mult * D), nn.Dropout(0.0), nn.Linear(mult * D, D)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class FluxTransformerBlock(nn.Module):
    def __init__(self, D, heads, head_dim):
        super().__init__()
        self.norm1, self.norm1_context = AdaLayerNormZero(D), AdaLayerNormZero(D)
        self.attn = Attention(D, heads, head_dim)
        self.norm2 = nn.LayerNorm(D, elementwise_affine=False, eps=1e-6)
        self.ff = FeedForward(D)
        self.norm2_context = nn.LayerNorm(D, elementwise_affine=False, eps=1e-6)
        self.ff_context = FeedForward(D)


class FluxSingleTransformerBlock(nn.Module):
    def __init__(self, D, heads, head_dim, mlp_ratio=4):
        super().__init__()
        self.norm = AdaLayerNormZeroSingle(D)
        self.proj_mlp = nn.Linear(D, mlp_ratio * D)
        self.act_mlp = nn.GELU(approximate="tanh")
        self.proj_out = nn.Linear(D + mlp_ratio * D, D)
        self.attn = Attention(D, heads, head_dim, pre_only=True)


def apply_rotary_emb(x, freqs_cis, use_real=True, use_real_unbind_dim=-1):
    cos, sin = freqs_cis
    cos, sin = cos[None, None].to(x.device), sin[None, None].to(x.device)
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rotated = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rotated.float() * sin).to(x.dtype)


class FluxPosEmbed(nn.Module):
    def __init__(self, theta, axes_dim):
        super().__init__()
        self.theta, self.axes_dim = theta, axes_dim

    def forward(self, ids):
        pos = ids.float()
        cos_out, sin_out = [], []
        for i, d in enumerate(self.axes_dim):
            freqs = 1.0 / (self.theta ** (torch.arange(0, d, 2, dtype=torch.float64)[: d // 2] / d))
            ang = torch.outer(pos[:, i].to(torch.float64), freqs)
            cos_out.append(ang.cos().repeat_interleave(2, dim=1).float())
            sin_out.append(ang.sin().repeat_interleave(2, dim=1).float())
        return torch.cat(cos_out, dim=-1), torch.cat(sin_out, dim=-1)


def _timesteps_256(t):
    half = 128
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    return torch.cat([emb[:, half:], emb[:, :half]], dim=-1)  # flip_sin_to_cos=True


class _MLP(nn.Module):  # TimestepEmbedding / PixArtAlphaTextProjection: Linear -> SiLU -> Linear
    def __init__(self, d_in, D):
        super().__init__()
        self.linear_1, self.linear_2 = nn.Linear(d_in, D), nn.Linear(D, D)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class CombinedTimestepGuidanceTextProjEmbeddings(nn.Module):
    def __init__(self, D, pooled_dim, guidance: bool):
        super().__init__()
        self.timestep_embedder = _MLP(256, D)
        if guidance:
            self.guidance_embedder = _MLP(256, D)
        self.text_embedder = _MLP(pooled_dim, D)
        self.has_guidance = guidance

    def forward(self, timestep, *rest):
        if self.has_guidance:
            guidance, pooled = rest
        else:
            (pooled,) = rest
        cond = self.timestep_embedder(_timesteps_256(timestep).to(pooled.dtype))
        if self.has_guidance:
            cond = cond + self.guidance_embedder(_timesteps_256(guidance).to(pooled.dtype))
        return cond + self.text_embedder(pooled)


class FluxTransformer2DModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        D, H, dh = cfg.inner_dim, cfg.num_attention_heads, cfg.attention_head_dim
        self.config = types.SimpleNamespace(in_channels=cfg.in_channels, guidance_embeds=cfg.guidance_embeds)
        self.gradient_checkpointing = False
        self.pos_embed = FluxPosEmbed(10000, cfg.axes_dims_rope)
        self.time_text_embed = CombinedTimestepGuidanceTextProjEmbeddings(D, cfg.pooled_projection_dim, cfg.guidance_embeds)
        self.context_embedder = nn.Linear(cfg.joint_attention_dim, D)
        self.x_embedder = nn.Linear(cfg.in_channels, D)
        self.transformer_blocks = nn.ModuleList([FluxTransformerBlock(D, H, dh) for _ in range(cfg.num_layers)])
        self.single_transformer_blocks = nn.ModuleList(
            [FluxSingleTransformerBlock(D, H, dh, cfg.mlp_ratio) for _ in range(cfg.num_single_layers)])
        self.norm_out = AdaLayerNormContinuous(D)
        self.proj_out = nn.Linear(D, cfg.in_channels)


class Transformer2DModelOutput:
    def __init__(self, sample):
        self.sample = sample


class FluxPipelineOutput:
    def __init__(self, images):
        self.images = images


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.16):
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None, **kwargs):
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed.")
    scheduler.set_timesteps(sigmas=sigmas, device=device, **kwargs)
    return scheduler.timesteps, len(scheduler.timesteps)


class FlowMatchEulerDiscreteScheduler:
    """diffusers 0.31.0 scheduler with the FLUX.1-dev scheduler_config.json."""

    order = 1

    def __init__(self):
        self.config = types.SimpleNamespace(num_train_timesteps=1000, shift=3.0, use_dynamic_shifting=True, base_shift=0.5,
                                            max_shift=1.15, base_image_seq_len=256, max_image_seq_len=4096)
        self._step_index = None

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None):
        sigmas = np.array(sigmas, dtype=np.float64)
        sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1) ** 1.0)  # time_shift(mu, 1.0, sigmas)
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        self.timesteps = (sigmas * self.config.num_train_timesteps).to(device=device)
        self.sigmas = torch.cat([sigmas, torch.zeros(1, device=sigmas.device)])
        self._step_index = None

    def step(self, model_output, timestep, sample, return_dict=True):
        if self._step_index is None:
            self._step_index = int((self.timesteps == timestep).nonzero()[0].item())
        sample = sample.to(torch.float32)
        sigma, sigma_next = self.sigmas[self._step_index], self.sigmas[self._step_index + 1]
        prev = sample + (sigma_next - sigma) * model_output
        prev = prev.to(model_output.dtype)
        self._step_index += 1
        return (prev,)


class FluxPipeline:
    """The slice of diffusers 0.31.0 FluxPipeline that generate.py / pipeline_tools.py / condition.py consume
    (SURVEY.md §8b "pipeline object", App. A.7).  Text encoders / VAE are out of scope: embeddings and condition
    latents are supplied pre-computed."""

    def __init__(self, transformer, dtype=torch.float32):
        self.transformer = transformer
        self.scheduler = FlowMatchEulerDiscreteScheduler()
        self.vae_scale_factor = 16
        self.default_sample_size = 64
        self.dtype, self.device = dtype, torch.device("cpu")
        self._execution_device = self.device
        self._interrupt = False
        self._joint_attention_kwargs = None

    joint_attention_kwargs = property(lambda self: self._joint_attention_kwargs)
    interrupt = property(lambda self: self._interrupt)

    def check_inputs(self, prompt, prompt_2, height, width, prompt_embeds=None, pooled_prompt_embeds=None,
                     callback_on_step_end_tensor_inputs=None, max_sequence_length=None):
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError("`height` and `width` have to be divisible by 8")
        if prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`.")

    def encode_prompt(self, prompt=None, prompt_2=None, prompt_embeds=None, pooled_prompt_embeds=None, device=None,
                      num_images_per_prompt=1, max_sequence_length=512, lora_scale=None):
        if prompt_embeds is None and isinstance(prompt, tuple):  # (prompt_embeds, pooled) smuggled through `prompt`
            prompt_embeds, pooled_prompt_embeds = prompt  # (model.py:585 passes batch["description"] to the text encoders)
        assert prompt_embeds is not None, "text encoders are out of scope: pass prompt_embeds"
        text_ids = torch.zeros(prompt_embeds.shape[1], 3).to(device=device, dtype=prompt_embeds.dtype)
        return prompt_embeds, pooled_prompt_embeds, text_ids

    @staticmethod
    def _prepare_latent_image_ids(batch_size, height, width, device, dtype):
        ids = torch.zeros(height // 2, width // 2, 3)
        ids[..., 1] = ids[..., 1] + torch.arange(height // 2)[:, None]
        ids[..., 2] = ids[..., 2] + torch.arange(width // 2)[None, :]
        h, w, c = ids.shape
        return ids.reshape(h * w, c).to(device=device, dtype=dtype)

    @staticmethod
    def _pack_latents(latents, batch_size, num_channels_latents, height, width):
        latents = latents.view(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
        latents = latents.permute(0, 2, 4, 1, 3, 5)
        return latents.reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)

    @staticmethod
    def _unpack_latents(latents, height, width, vae_scale_factor):
        batch_size, num_patches, channels = latents.shape
        height, width = height // vae_scale_factor, width // vae_scale_factor
        latents = latents.view(batch_size, height, width, channels // 4, 2, 2)
        latents = latents.permute(0, 3, 1, 4, 2, 5)
        return latents.reshape(batch_size, channels // (2 * 2), height * 2, width * 2)

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        height = 2 * (int(height) // self.vae_scale_factor)
        width = 2 * (int(width) // self.vae_scale_factor)
        if latents is not None:
            ids = self._prepare_latent_image_ids(batch_size, height, width, device, dtype)
            return latents.to(device=device, dtype=dtype), ids
        shape = (batch_size, num_channels_latents, height, width)
        latents = torch.randn(shape, generator=generator, device=device, dtype=dtype)  # randn_tensor
        latents = self._pack_latents(latents, batch_size, num_channels_latents, height, width)
        return latents, self._prepare_latent_image_ids(batch_size, height, width, device, dtype)

    def set_adapters(self, name):
        self.adapter = name

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        yield types.SimpleNamespace(update=lambda: None)

    def maybe_free_model_hooks(self):
        pass


class _PrecomputedVae:
    """`encode_images` (pipeline_tools.py:7-14) with the VAE factored out: `images` already are latents z, so that
    (z - shift) * scale, packing and the id fallback (:15-29) execute for real."""

    def __init__(self, shift=0.1159, scale=0.3611):
        self.config = types.SimpleNamespace(shift_factor=shift, scaling_factor=scale)

    def encode(self, z):
        return types.SimpleNamespace(latent_dist=types.SimpleNamespace(sample=lambda: z))


# ------------------------------------------------------------------------------------------------------------------
# stub installation + import of the reference tree
# ------------------------------------------------------------------------------------------------------------------
_displaced: Dict[str, object] = {}  # stand-in name -> what sys.modules held before (None = nothing)


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so `import a.b` resolves through sys.modules
    _displaced.setdefault(name, sys.modules.get(name))
    sys.modules[name] = m
    return m


_installed = False


def uninstall_stubs():
    """Take the stand-ins (and the reference modules imported on top of them) out of sys.modules again and put back what
    they displaced.  Libraries that probe for optional packages with importlib.util.find_spec (transformers: accelerate,
    peft) trip over stand-ins without a spec, so a test module that used the harness removes them when it is done."""
    global _installed
    for name, prev in _displaced.items():
        if prev is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = prev
    _displaced.clear()
    for name in [n for n in sys.modules if n == PKG or n.startswith(PKG + ".")]:
        del sys.modules[name]
    _installed = False


def install_stubs():
    """Register stand-ins for the third-party imports at the top of the reference files (block.py:3, transformer.py:2-14,
    generate.py:3-13, lora_controller.py:1, pipeline_tools.py:1-3, condition.py:3, model.py:1-14)."""
    global _installed
    if _installed:
        return
    from oracle import cs3_dgf as C

    log = logging.getLogger("loongx_reference")
    _module("diffusers")
    _module("diffusers.pipelines", FluxPipeline=FluxPipeline)
    _module("diffusers.pipelines.flux")
    _module("diffusers.pipelines.flux.pipeline_flux", FluxPipelineOutput=FluxPipelineOutput, calculate_shift=calculate_shift,
            retrieve_timesteps=retrieve_timesteps, np=np, logger=log)
    _module("diffusers.utils", logging=logging)
    _module("diffusers.models")
    _module("diffusers.models.attention_processor", Attention=Attention, F=F)
    _module("diffusers.models.embeddings", apply_rotary_emb=apply_rotary_emb)
    _module("diffusers.models.transformers")
    _module("diffusers.models.transformers.transformer_flux", FluxTransformer2DModel=FluxTransformer2DModel,
            Transformer2DModelOutput=Transformer2DModelOutput, USE_PEFT_BACKEND=True,
            scale_lora_layers=lambda model, weight: None if weight == 1.0 else _scale_all(model, weight),
            unscale_lora_layers=lambda model, weight=None: None if weight in (None, 1.0) else _scale_all(model, 1 / weight),
            logger=log)
    _module("accelerate")
    _module("accelerate.utils", is_torch_version=lambda op, v: True)
    _module("peft", LoraConfig=dict, get_peft_model_state_dict=lambda m: {})
    _module("peft.tuners")
    _module("peft.tuners.tuners_utils", BaseTunerLayer=BaseTunerLayer)
    _module("lightning", LightningModule=nn.Module)
    _module("prodigyopt")

    def S4Model(d_input, d_model, d_output, n_blocks, n, l_max, **kw):  # s4torch signature used at model.py:31-38
        return C.S4Model(d_input, d_model, d_output, n_blocks, n, l_max)

    _module("s4torch", S4Model=S4Model)
    ref = types.ModuleType(PKG)
    ref.__path__ = [os.path.join(REF_ROOT, "src")]
    sys.modules[PKG] = ref
    _installed = True


def _scale_all(model, w):
    for m in model.modules():
        if isinstance(m, BaseTunerLayer):
            m.scale_layer(w)


def ref_module(name: str):
    """import /root/reference/src/<name> (e.g. "flux.block", "train.model") with the stubs in place."""
    if not available():
        raise RuntimeError("reference tree not present at " + REF_ROOT)
    install_stubs()
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # /root/reference is read-only
    try:
        return importlib.import_module(f"{PKG}.{name}")
    finally:
        sys.dont_write_bytecode = dont


# ------------------------------------------------------------------------------------------------------------------
# oracle flat params <-> stand-in modules
# ------------------------------------------------------------------------------------------------------------------
def build_transformer(P: Dict[str, torch.Tensor], cfg, dtype=torch.float32) -> FluxTransformer2DModel:
    """Stand-in FluxTransformer2DModel holding exactly the oracle's parameters; LoRA targets (oracle.flux_dit
    .lora_target_names = regex of train/config/seed_512.yaml:38) are wrapped in LoraLinear like peft's add_adapter."""
    from oracle import flux_dit as O

    model = FluxTransformer2DModel(cfg)
    for name in O.lora_target_names(cfg):
        if name + ".lora_A.weight" not in P:
            continue
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        base = parent[int(leaf)] if leaf.isdigit() else getattr(parent, leaf)
        wrapped = LoraLinear(base, cfg.lora_rank, cfg.lora_alpha)
        if leaf.isdigit():
            parent[int(leaf)] = wrapped
        else:
            setattr(parent, leaf, wrapped)
    sd = {}
    for k, v in P.items():
        stem, _, kind = k.rpartition(".")
        if stem.endswith(".lora_A") or stem.endswith(".lora_B"):
            sd[f"{stem}.default.weight"] = v
        elif stem + ".lora_A.weight" in P:
            sd[f"{stem}.base_layer.{kind}"] = v
        else:
            sd[k] = v
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing[:5], unexpected[:5])
    return model.to(dtype).eval()


def copy_module_params(dst: nn.Module, src: nn.Module):
    """Copy parameters/buffers between two structurally equal modules by position (names differ between the reference
    encoders and the oracle's)."""
    d, s = list(dst.state_dict().items()), list(src.state_dict().items())
    assert len(d) == len(s), (len(d), len(s))
    with torch.no_grad():
        for (kd, vd), (ks, vs) in zip(d, s):
            assert vd.shape == vs.shape, (kd, ks, vd.shape, vs.shape)
            vd.copy_(vs)
