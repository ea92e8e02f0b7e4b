"""Pipeline / transformer host objects with the attribute surface the reference's `generate()` and `OminiModel`
consume from diffusers' FluxPipeline / FluxTransformer2DModel ("pipeline object consumed by generate()").

They own the packed native weights and cached plans; every tensor operation they perform is either a native kernel or
pure layout plumbing (views / copies).  The VAE and the text encoders are optional attachments
(`attach_vae`, `attach_text_encoders`): without them prompt embeddings and latents come in pre-computed and
`output_type="latent"` is the supported output.
"""
from __future__ import annotations

import contextlib
from typing import Dict, List, Optional, Tuple

import torch

from .config import FluxConfig
from .dit import DitPlan, DitWeights, pack_latents, random_params, unpack_latents
from .sampler import FlowMatchEulerDiscreteScheduler, latent_image_ids


class _Cfg(dict):
    """dict with attribute access (diffusers FrozenDict-like)."""

    __getattr__ = dict.__getitem__


class _AttnHandle:
    """Stand-in for a diffusers `Attention` module: generate() only sets / deletes `.c_factor` on it
    (generate.py:90-94, 385-389) and attn_forward() reads it (block.py:121-128)."""

    def __init__(self, heads: int):
        self.heads = heads


class _BlockHandle:
    def __init__(self, transformer: "NativeFluxTransformer", index: int, single: bool):
        self.transformer, self.index, self.single = transformer, index, single
        self.attn = _AttnHandle(transformer.cfg.num_attention_heads)
        self.attn.block = self  # src.flux.block.attn_forward(attn, ...) finds its weights through the block


class NativeFluxTransformer:
    """FluxTransformer2DModel-shaped owner of the native DiT weights."""

    def __init__(self, cfg: FluxConfig, params: Optional[Dict[str, torch.Tensor]] = None, device="cuda", seed: int = 1234,
                 consume_params: bool = False):
        cfg.validate()
        self.cfg = cfg
        self.device = torch.device(device)
        self.dtype = torch.bfloat16
        if params is None:
            params = random_params(cfg, self.device, seed=seed)
            consume_params = True
        self.weights = DitWeights(params, cfg, self.device, consume=consume_params)
        self.config = _Cfg(in_channels=cfg.in_channels, guidance_embeds=cfg.guidance_embeds,
                           num_layers=cfg.num_layers, num_single_layers=cfg.num_single_layers,
                           attention_head_dim=cfg.attention_head_dim, num_attention_heads=cfg.num_attention_heads,
                           joint_attention_dim=cfg.joint_attention_dim, pooled_projection_dim=cfg.pooled_projection_dim,
                           axes_dims_rope=cfg.axes_dims_rope)
        self.training = False
        self.gradient_checkpointing = False
        self.transformer_blocks = [_BlockHandle(self, i, False) for i in range(cfg.num_layers)]
        self.single_transformer_blocks = [_BlockHandle(self, i, True) for i in range(cfg.num_single_layers)]
        self._plans: Dict[tuple, DitPlan] = {}

    # -- checkpoints --------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, flux_path: str, device="cuda", lora_rank: int = 4, lora_alpha: float = 4.0, seed: int = 1234):
        """diffusers-format FLUX transformer directory -> native weights; fresh LoRA factors like
        `transformer.add_adapter(LoraConfig(init_lora_weights="gaussian"))` (model.py:519): A ~ N(0, 1/r), B = 0."""
        from .checkpoint import read_diffusers_transformer

        cfg, P = read_diffusers_transformer(flux_path, device=device, lora_rank=lora_rank, lora_alpha=lora_alpha)
        init_lora_factors(P, cfg, device, seed)
        return cls(cfg, P, device=device, consume_params=True)

    def load_params(self, params: Dict[str, torch.Tensor]) -> None:
        """Replace every weight (base + LoRA factors) from a flat diffusers-named dict; drops the cached plans."""
        from .checkpoint import check_transformer_params

        check_transformer_params(params, self.cfg)
        self._plans.clear()
        self.weights = None  # free the old panels before packing the new ones
        self.weights = DitWeights(dict(params), self.cfg, self.device, consume=True)

    def load_lora_factors(self, lora: Dict[str, torch.Tensor]) -> int:
        """Overwrite LoRA factors in place (`<module>.lora_A.weight` / `.lora_B.weight`) and re-merge the affected
        panels with the native merge kernel; returns the number of modules updated.  A file whose rank differs from the
        constructed one re-packs the weight set at the file's rank (peft's load_lora_weights accepts any rank)."""
        from .dit import PackedLinear
        from .train import LoraFactor, lora_shared

        targets = {}
        for panel in self.weights.named.values():
            if isinstance(panel, PackedLinear):
                for (name, row0, rows, A, Bw) in panel.lora:
                    targets[name] = (panel, row0, rows, A, Bw)
        mods = {k.rsplit(".lora_", 1)[0] for k in lora}
        unknown = mods - set(targets)
        if unknown:
            raise KeyError(f"LoRA factors for modules that are not LoRA targets here: {sorted(unknown)[:4]}")
        for name in mods:
            ka, kb = name + ".lora_A.weight", name + ".lora_B.weight"
            if ka not in lora or kb not in lora:
                raise KeyError(f"{name}: a LoRA checkpoint must hold both lora_A and lora_B (found only "
                               f"{'lora_A' if ka in lora else 'lora_B'})")
            if lora[ka].shape[0] != lora[kb].shape[1]:
                raise ValueError(f"{name}: lora_A {tuple(lora[ka].shape)} and lora_B {tuple(lora[kb].shape)} disagree on the rank")
        ranks = {lora[name + ".lora_A.weight"].shape[0] for name in mods}
        if mods and ranks != {self.cfg.lora_rank}:
            if len(ranks) != 1:
                raise ValueError(f"LoRA checkpoint mixes ranks {sorted(ranks)}")
            return self._repack_with_lora(lora, ranks.pop())
        shared = lora_shared(self.weights)
        scale = getattr(self.weights, "lora_scale", 1.0)
        for name in mods:
            panel, row0, rows, A, Bw = targets[name]
            a, b = lora[name + ".lora_A.weight"], lora[name + ".lora_B.weight"]
            if a.shape != A.shape or b.shape != Bw.shape:
                raise ValueError(f"{name}: LoRA factors {tuple(a.shape)} / {tuple(b.shape)} do not match "
                                 f"rank-{A.shape[0]} factors {tuple(A.shape)} / {tuple(Bw.shape)}")
            with torch.no_grad():
                shared[name].A.copy_(a.to(device=A.device, dtype=A.dtype))
                shared[name].B.copy_(b.to(device=A.device, dtype=A.dtype))
            LoraFactor(name, panel, row0, rows, shared[name]).remerge(scale)
        return len(mods)

    def _repack_with_lora(self, lora: Dict[str, torch.Tensor], rank: int) -> int:
        """Rank change: export the flat parameter dict, swap the factors (targets missing from the file get fresh
        zero-B factors of the new rank), re-pack.  alpha keeps its constructed value, so the scaling becomes alpha / rank
        (what peft computes from the LoraConfig when the adapter is re-created at another rank)."""
        from .config import lora_targets

        P = self.weights.export_params()
        for k in [k for k in P if ".lora_" in k]:
            del P[k]
        self.cfg.lora_rank = int(rank)
        for name in lora_targets(self.cfg):
            ka, kb = name + ".lora_A.weight", name + ".lora_B.weight"
            if ka in lora:
                P[ka], P[kb] = lora[ka].to(self.device).float(), lora[kb].to(self.device).float()
        init_lora_factors(P, self.cfg, self.device)
        self.load_params(P)
        return len({k.rsplit(".lora_", 1)[0] for k in lora})

    def set_lora_scale(self, scale: float) -> None:
        """`joint_attention_kwargs["scale"]` (transformer.py:73-83: peft `scale_lora_layers(self, lora_scale)` around the
        forward): re-merge every LoRA-carrying panel as W + scale * (alpha / r) B A with the native merge kernel.  A no-op
        while the scale does not change; the reference's un-scaling at the end of the forward (transformer.py:246-248)
        corresponds to calling this with 1.0 again.  The scale in force lives on the weight set (train.set_lora_scale)."""
        from .train import set_lora_scale

        self.weights.lora_inner = float(scale)
        set_lora_scale(self.weights, float(scale) * getattr(self.weights, "lora_outer", 1.0))

    def set_lora_outer(self, factor: float) -> None:
        """The multiplier the reference's context managers put on peft's `scaling` from OUTSIDE a forward
        (lora_controller.py: `enable_lora(..., False)` = 0, `set_lora_scale(..., s)` = s): like there it composes with the
        per-forward `joint_attention_kwargs["scale"]` (peft's scale_lora_layers multiplies), so the scale in force is
        inner * outer."""
        from .train import set_lora_scale

        self.weights.lora_outer = float(factor)
        set_lora_scale(self.weights, getattr(self.weights, "lora_inner", 1.0) * float(factor))

    def remerge_lora(self) -> None:
        """Rebuild every merged panel from the current LoRA factors (after an optimizer step or an in-place edit of
        `lora_parameters()`), at the LoRA scale in force."""
        from .dit import PackedLinear
        from .train import LoraFactor, lora_shared

        shared = lora_shared(self.weights)
        scale = getattr(self.weights, "lora_scale", 1.0)
        for panel in self.weights.named.values():
            if isinstance(panel, PackedLinear):
                for (name, row0, rows, _A, _B) in panel.lora:
                    LoraFactor(name, panel, row0, rows, shared[name]).remerge(scale)

    def lora_parameters(self):
        """model.py:513-524: the LoRA factors as nn.Parameters (one stable set per weight set)."""
        from .train import lora_parameters

        return lora_parameters(self.weights)

    def to(self, *args, **kwargs):
        """nn.Module.to for the calls the reference makes (inference.py:55-56, model.py:398): the native weights live in
        HBM in bf16 and stay there; a CUDA target is accepted, anything else raises."""
        device = kwargs.get("device", None)
        dtype = kwargs.get("dtype", None)
        for a in args:
            if isinstance(a, torch.dtype):
                dtype = a
            elif a is not None:
                device = a
        if device is not None and torch.device(device).type != "cuda":
            raise NotImplementedError("the native DiT weights live in HBM: there is no CPU representation to move to")
        if device is not None and torch.device(device).index not in (None, self.device.index):
            raise NotImplementedError(f"the weights were packed on {self.device}; re-create the model on {device}")
        if dtype not in (None, torch.bfloat16, torch.float32):
            raise NotImplementedError(f"dtype {dtype}: the native DiT computes in bf16 with fp32 accumulation")
        return self

    # -- nn.Module-like surface --------------------------------------------------------------------------------
    def named_modules(self):
        yield "", self
        for i, b in enumerate(self.transformer_blocks):
            yield f"transformer_blocks.{i}", b
            yield f"transformer_blocks.{i}.attn", b.attn
        for i, b in enumerate(self.single_transformer_blocks):
            yield f"single_transformer_blocks.{i}", b
            yield f"single_transformer_blocks.{i}.attn", b.attn

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def c_factor(self) -> Optional[float]:
        """The value generate() installed on every `*.attn` (None when condition_scale == 1)."""
        for b in self.transformer_blocks + self.single_transformer_blocks:
            cf = getattr(b.attn, "c_factor", None)
            if cf is not None:
                return float(cf.reshape(-1)[0]) if isinstance(cf, torch.Tensor) else float(cf)
        return None

    # -- plans -------------------------------------------------------------------------------------------------
    def plan(self, B: int, n_txt: int, n_img: int, n_cond: int, T: int, model_config: Optional[dict],
             c_factor: Optional[float], cache_cond: bool = False) -> DitPlan:
        mc = model_config or {}
        key = (cache_cond, B, n_txt, n_img, n_cond, T, bool(mc.get("latent_lora", False)), bool(mc.get("union_cond_attn", True)),
               bool(mc.get("independent_condition", False)), bool(mc.get("add_cond_attn", False)), c_factor)
        pl = self._plans.get(key)
        if pl is None:
            if len(self._plans) >= 4:  # bounded cache: plans own large activation buffers
                self._plans.pop(next(iter(self._plans)))
            pl = DitPlan(self.weights, B, n_txt, n_img, n_cond, T=T, model_config=mc, c_factor=c_factor, cache_cond=cache_cond)
            self._plans[key] = pl
        return pl


def init_lora_factors(P: Dict[str, torch.Tensor], cfg: FluxConfig, device, seed: int = 1234, b_std: float = 0.0) -> None:
    """peft 'gaussian' initialisation of every LoRA target that has no factors yet (in place)."""
    from .config import lora_targets

    if cfg.lora_rank <= 0:
        return
    g = torch.Generator(device=device).manual_seed(seed)
    for name in lora_targets(cfg):
        if name + ".lora_A.weight" in P:
            continue
        o, i = P[name + ".weight"].shape
        P[name + ".lora_A.weight"] = torch.randn(cfg.lora_rank, i, generator=g, device=device) / cfg.lora_rank
        P[name + ".lora_B.weight"] = (torch.randn(o, cfg.lora_rank, generator=g, device=device) * b_std) if b_std > 0 else \
            torch.zeros(o, cfg.lora_rank, device=device)


class _Scheduler(FlowMatchEulerDiscreteScheduler):
    pass


class NativeFluxPipeline:
    """FluxPipeline-shaped object for generate() (attribute list in SURVEY.md §8b)."""

    def __init__(self, transformer: NativeFluxTransformer, scheduler: Optional[FlowMatchEulerDiscreteScheduler] = None):
        self.transformer = transformer
        self.scheduler = scheduler or _Scheduler()
        self.vae = None
        self.text_encoder = None
        self.text_encoder_2 = None
        self.tokenizer = None
        self.tokenizer_2 = None
        self.image_processor = None
        self.vae_scale_factor = 16  # diffusers 0.31.0
        self.default_sample_size = 64
        self._guidance_scale = 3.5
        self._joint_attention_kwargs = None
        self._interrupt = False
        self._num_timesteps = 0

    # properties generate() reads
    @property
    def device(self):
        return self.transformer.device

    @property
    def dtype(self):
        return self.transformer.dtype

    @property
    def _execution_device(self):
        return self.transformer.device

    @property
    def joint_attention_kwargs(self):
        return self._joint_attention_kwargs

    @property
    def interrupt(self):
        return self._interrupt

    @property
    def guidance_scale(self):
        return self._guidance_scale

    # FluxPipeline methods generate() calls
    def check_inputs(self, prompt, prompt_2, height, width, prompt_embeds=None, pooled_prompt_embeds=None,
                     callback_on_step_end_tensor_inputs=None, max_sequence_length=None):
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        if prompt is not None and prompt_embeds is not None:
            raise ValueError("Cannot forward both `prompt` and `prompt_embeds`.")
        if prompt is None and prompt_embeds is None:
            raise ValueError("Provide either `prompt` or `prompt_embeds`.")
        if prompt_embeds is not None and pooled_prompt_embeds is None:
            raise ValueError("If `prompt_embeds` are provided, `pooled_prompt_embeds` also have to be passed.")
        if max_sequence_length is not None and max_sequence_length > 512:
            raise ValueError(f"`max_sequence_length` cannot be greater than 512 but is {max_sequence_length}")

    def encode_prompt(self, prompt=None, prompt_2=None, device=None, num_images_per_prompt: int = 1, prompt_embeds=None,
                      pooled_prompt_embeds=None, max_sequence_length: int = 512, lora_scale=None):
        if prompt_embeds is None:
            if self.text_encoder is None or self.text_encoder_2 is None or self.tokenizer is None or self.tokenizer_2 is None:
                raise NotImplementedError(
                    "no text encoders attached (pipeline.attach_text_encoders, SURVEY.md §8f.4): pass prompt_embeds and "
                    "pooled_prompt_embeds")
            from .text import tokenize

            # FluxPipeline.encode_prompt: CLIP-L pooled output of `prompt` (77 tokens), T5 hidden states of `prompt_2`
            prompts = [prompt] if isinstance(prompt, str) else list(prompt)
            prompts_2 = prompts if prompt_2 is None else ([prompt_2] if isinstance(prompt_2, str) else list(prompt_2))
            pooled_prompt_embeds = self.text_encoder(tokenize(self.tokenizer, prompts, 77)).pooler_output
            prompt_embeds = self.text_encoder_2(tokenize(self.tokenizer_2, prompts_2, max_sequence_length))[0]
        device = device or self._execution_device
        pe = prompt_embeds.to(device=device, dtype=self.dtype)
        po = pooled_prompt_embeds.to(device=device, dtype=self.dtype)
        if num_images_per_prompt != 1:
            pe = pe.repeat_interleave(num_images_per_prompt, dim=0)
            po = po.repeat_interleave(num_images_per_prompt, dim=0)
        text_ids = torch.zeros(pe.shape[1], 3, device=device, dtype=self.dtype)
        return pe, po, text_ids

    @staticmethod
    def _prepare_latent_image_ids(batch_size, height, width, device, dtype):
        return torch.from_numpy(latent_image_ids(height // 2, width // 2)).to(device=device, dtype=dtype)

    @staticmethod
    def _pack_latents(latents, batch_size=None, num_channels_latents=None, height=None, width=None):
        return pack_latents(latents.contiguous())

    @staticmethod
    def _unpack_latents(latents, height, width, vae_scale_factor):
        return unpack_latents(latents.contiguous(), height, width, vae_scale_factor)

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        height = 2 * (int(height) // self.vae_scale_factor)
        width = 2 * (int(width) // self.vae_scale_factor)
        shape = (batch_size, num_channels_latents, height, width)
        ids = self._prepare_latent_image_ids(batch_size, height, width, device, dtype)
        if latents is not None:
            return latents.to(device=device, dtype=dtype), ids
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an "
                             f"effective batch size of {batch_size}.")
        if isinstance(generator, list):  # diffusers randn_tensor: sample i comes from generator[i]
            noise = torch.cat([torch.randn((1,) + shape[1:], generator=g, device=g.device, dtype=dtype).to(device)
                               for g in generator], dim=0)
        else:
            gen_dev = generator.device if generator is not None else torch.device(device)
            noise = torch.randn(shape, generator=generator, device=gen_dev, dtype=dtype).to(device)
        return self._pack_latents(noise), ids

    def set_adapters(self, *args, **kwargs):  # LoRA adapters are merged into the cond row group at load
        return None

    def to(self, *args, **kwargs):
        """FluxPipeline.to(...) as the reference calls it (inference.py:56 `model.flux_pipe.to("cuda")`, model.py:398
        `.to(dtype=dtype).to(device)`): every component already lives on the GPU."""
        self.transformer.to(*args, **kwargs)
        return self

    def attach_text_encoders(self, source=None, tokenizers=None, clip=None, t5=None):
        """Give the pipeline `text_encoder` (CLIP-L), `text_encoder_2` (T5-XXL) and their tokenizers so
        that `generate(prompt=...)` / `prepare_text_input` work.  `source`: a FLUX checkpoint directory (text_encoder/,
        text_encoder_2/, tokenizer/, tokenizer_2/); or pass already-built encoders and tokenizer callables."""
        from .text import load_text_encoders, load_tokenizers

        if source is not None:
            clip, t5 = load_text_encoders(source, self.device)
            tokenizers = tokenizers or load_tokenizers(source)
        if clip is None or t5 is None or tokenizers is None:
            raise ValueError("attach_text_encoders needs a checkpoint directory, or clip=, t5= and tokenizers=")
        self.text_encoder, self.text_encoder_2 = clip, t5
        self.tokenizer, self.tokenizer_2 = tokenizers
        return clip, t5

    def attach_vae(self, source=None, seed: int = 1234):
        """Give the pipeline its `vae` + `image_processor`.  `source`: a FLUX checkpoint directory
        (reads <dir>/vae), a flat diffusers-named parameter dict, a `VaeWeights`, or None for seeded synthetic
        parameters of FLUX.1-dev's VAE architecture."""
        from .vae import ImageProcessor, NativeVae, VaeConfig, VaeWeights, synthetic_params

        if isinstance(source, VaeWeights):
            w = source
        elif isinstance(source, str):
            w = VaeWeights.from_pretrained(source, self.device)
        else:
            cfg = VaeConfig()
            w = VaeWeights(cfg, source if source is not None else synthetic_params(cfg, seed), self.device)
        self.vae = NativeVae(w)
        self.vae_scale_factor = 2 ** len(w.cfg.block_out_channels)  # diffusers 0.31.0 FluxPipeline: 16
        self.image_processor = ImageProcessor(self.vae_scale_factor)
        return self.vae

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        class _Bar:
            def update(self, n=1):
                pass

        yield _Bar()

    def maybe_free_model_hooks(self):
        return None


class FluxPipelineOutput:
    def __init__(self, images):
        self.images = images
