"""Drop-in for the reference's src/train/model.py: OminiModel + the CS3 / DGF modules (model.py:16-1035).

Constructor signature and the attribute surface consumed by generate() (model.py:377-462) are kept;
the sub-modules are the native-kernel shells from loongx_b200.cs3.  `step()` / `training_step()` (model.py:560-729) run
the native forward + backward of loongx_b200.train; the Lightning harness itself and checkpoint serialisation
(model.py:780-943) are outside this build's scope.
"""
from typing import Optional

import torch
import torch.nn as nn

from loongx_b200 import cs3
from loongx_b200.config import FluxConfig
from loongx_b200.cs3 import DUAN, EEGEncoder, FeaturePyramidPooling, FNIRSEncoder, MotionEncoder, PPGEncoder  # noqa: F401
from loongx_b200.pipeline import NativeFluxPipeline, NativeFluxTransformer


class _StagedPipeline:
    """`model.flux_pipe` between `OminiModel(device="cpu")` and `model.to("cuda")` (inference.py:35-56): the native
    engine has no host representation, so nothing is built yet; `.to("cuda")` here materialises the model as well."""

    def __init__(self, owner: "OminiModel"):
        self._owner = owner

    def to(self, *args, **kwargs):
        self._owner.to(*args, **kwargs)
        return self._owner.flux_pipe

    def __getattr__(self, name):
        raise RuntimeError(f"flux_pipe.{name}: the model was created with device='cpu' and is only staged; call "
                           "model.to('cuda') first (the native DiT has no CPU path)")


def _parse_to(args, kwargs):
    device, dtype = kwargs.get("device"), kwargs.get("dtype")
    for a in args:
        if isinstance(a, torch.dtype):
            dtype = a
        elif isinstance(a, torch.Tensor):
            device, dtype = a.device, a.dtype
        elif a is not None:
            device = a
    return (torch.device(device) if device is not None else None), dtype


class OminiModel(nn.Module):
    def __init__(self, flux_pipe_id, lora_path: str = None, lora_config: dict = None, device: str = "cuda",
                 dtype: torch.dtype = torch.bfloat16, model_config: dict = {}, optimizer_config: dict = None,
                 gradient_checkpointing: bool = False, use_brain_condition: bool = True, fuse_flag: bool = True,
                 seed: int = 1234):
        """`flux_pipe_id`: a FluxConfig (random-init weights of that architecture, seeded) or the string "synthetic"
        (FLUX.1-dev geometry), or a diffusers-format FLUX directory (transformer/ and, when present, vae/, text encoders).

        `dtype`: torch.bfloat16 or torch.float32 (train/config/seed_512.yaml:2 says "float32", the only dtype the
        reference's CS3 / DGF modules work in).  Either way the DiT computes in bf16 with fp32 accumulation (the native
        engine's one precision) and CS3 / DGF in float32; the value is kept as `_dtype` like model.py:394.

        `device`: "cuda[:i]" builds the native engine right away.  "cpu" (what inference.py:35-41 passes) STAGES the
        model: the CS3 / DGF modules are created on the host, `load_lora` / `load_state_dict` are recorded, and
        `model.to("cuda")` (inference.py:55) builds the native engine and replays them.  Nothing computes on the CPU."""
        super().__init__()
        if dtype not in (torch.bfloat16, torch.float32):
            raise NotImplementedError(f"dtype {dtype}: the native DiT computes in bf16 (fp32 accumulate); CS3/DGF run in float32")
        import os

        r = int((lora_config or {}).get("r", 4))
        alpha = float((lora_config or {}).get("lora_alpha", 4.0))
        pretrained = None
        if isinstance(flux_pipe_id, FluxConfig):
            cfg = flux_pipe_id
        elif flux_pipe_id == "synthetic":
            cfg = FluxConfig()
        elif isinstance(flux_pipe_id, str) and os.path.isdir(flux_pipe_id):
            pretrained, cfg = flux_pipe_id, None  # diffusers-format directory (FluxPipeline.from_pretrained, model.py:398)
        else:
            raise FileNotFoundError(f"{flux_pipe_id!r}: expected a FluxConfig, 'synthetic' or a diffusers-format FLUX directory")
        if cfg is not None and lora_config is not None:
            cfg.lora_rank, cfg.lora_alpha = r, alpha
        if lora_path:
            raise NotImplementedError  # model.py:517 raises as well
        self.model_config = model_config
        self.optimizer_config = optimizer_config
        self._dtype = dtype
        self._device = torch.device(device)
        self._build_args = dict(cfg=cfg, pretrained=pretrained, r=r, alpha=alpha, seed=seed,
                                gradient_checkpointing=gradient_checkpointing)
        self._pending = []  # load_lora / load_state_dict calls recorded while staged
        torch.manual_seed(seed)
        self.fuse_flag = fuse_flag
        self.use_brain_condition = use_brain_condition
        self.eeg_fixed_length, self.fnirs_fixed_length, self.ppg_fixed_length, self.motion_fixed_length = 4096, 512, 256, 128

        self.transformer = None
        self.flux_pipe = _StagedPipeline(self)
        if self._device.type == "cuda":
            self._build_native(self._device)

        f32 = dict(device=device, dtype=torch.float32)
        self.fusion1 = nn.Sequential(nn.Linear(512 * 2, 512)).to(**f32)
        self.fusion2 = nn.Sequential(nn.Linear(768 + 768, 768)).to(**f32)
        self.duan_norm1 = DUAN(channels=512, device=device)
        self.duan_norm2 = DUAN(channels=1, device=device)
        self.fusion3 = nn.Sequential(nn.Linear(512 * 2, 512)).to(**f32)
        self.fusion4 = nn.Sequential(nn.Linear(768 * 2, 768)).to(**f32)
        self.duan_norm_prompt = DUAN(channels=512, device=device)
        self.duan_norm_pooled = DUAN(channels=1, device=device)
        self.eeg_projection = EEGEncoder(device=device)
        self.ppg_projection = PPGEncoder(device=device)
        self.fnirs_projection = FNIRSEncoder(device=device)
        self.motion_projection = MotionEncoder(device=device)
        self.eval()

    # ---- device handling (inference.py:35-58) ----------------------------------------------------------------------
    def _build_native(self, device: torch.device) -> None:
        """FluxPipeline.from_pretrained(...).to(dtype).to(device) (model.py:397-400) for the native engine."""
        import os

        a = self._build_args
        pretrained = a["pretrained"]
        if pretrained is not None:
            self.transformer = NativeFluxTransformer.from_pretrained(pretrained, device=device, lora_rank=a["r"],
                                                                     lora_alpha=a["alpha"], seed=a["seed"])
        else:
            self.transformer = NativeFluxTransformer(a["cfg"], device=device, seed=a["seed"])
        self.transformer.gradient_checkpointing = a["gradient_checkpointing"]
        self.transformer.train(self.training)
        self.flux_pipe = NativeFluxPipeline(self.transformer)
        if pretrained is not None and os.path.isdir(os.path.join(pretrained, "vae")):
            self.flux_pipe.attach_vae(pretrained)  # FluxPipeline.from_pretrained loads the VAE too (model.py:398-400)
        if pretrained is not None and all(os.path.isdir(os.path.join(pretrained, d)) for d in
                                          ("text_encoder", "text_encoder_2", "tokenizer", "tokenizer_2")):
            self.flux_pipe.attach_text_encoders(pretrained)  # ... and both text encoders with their tokenizers

    @property
    def device(self):
        return self._device

    @property
    def staged(self) -> bool:
        return self.transformer is None

    def to(self, *args, **kwargs):
        """nn.Module.to restricted to what the native engine can honour: a CUDA target materialises a staged model
        (replaying the recorded checkpoint loads, inference.py:43-55) and moves the CS3 / DGF modules; dtype requests
        other than bf16 / fp32 raise; parameters keep their dtypes (CS3 / DGF fp32 + complex64, DiT bf16)."""
        device, dtype = _parse_to(args, kwargs)
        if dtype not in (None, torch.bfloat16, torch.float32):
            raise NotImplementedError(f"dtype {dtype}: the native DiT computes in bf16; CS3 / DGF run in float32")
        if dtype is not None:
            self._dtype = dtype
        if device is None:
            return self
        if device.type != "cuda":
            if self.staged:
                return self
            raise NotImplementedError("the native DiT weights live in HBM: a materialised model cannot move to the CPU")
        if not torch.cuda.is_available():
            raise RuntimeError("OminiModel.to('cuda'): no CUDA device — the native engine has no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if self.staged:
            self._build_native(device)
        else:
            self.transformer.to(device)
        super().to(device=device)  # CS3 / DGF modules (fp32 parameters, complex64 S4 factors)
        self._device = device
        pending, self._pending = self._pending, []
        for kind, payload in pending:
            if kind == "lora":
                self.load_lora(payload)
            else:
                self._load_transformer_state(payload)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device) if isinstance(device, int) else (device or "cuda"))

    def train(self, mode: bool = True):
        super().train(mode)
        if getattr(self, "transformer", None) is not None:
            self.transformer.train(mode)
        return self

    # ---- checkpoints (model.py:464-477, 526-531; inference.py:43-53) ------------------------------------------------
    def load_lora(self, checkpoint_path: str):
        """peft LoRA weights written by `save_lora` / FluxPipeline.save_lora_weights -> factors + native re-merge
        (model.py:464-477: also puts the transformer in eval mode)."""
        if self.staged:
            self._pending.append(("lora", checkpoint_path))
            return self
        from loongx_b200.checkpoint import read_peft_lora

        self.transformer.load_lora_factors(read_peft_lora(checkpoint_path, device=self.device))
        self.transformer.eval()
        return self

    def save_lora(self, path: str):
        """model.py:526-531: `pytorch_lora_weights.safetensors` in the layout of FluxPipeline.save_lora_weights."""
        import os

        from safetensors.torch import save_file

        os.makedirs(path, exist_ok=True)
        P = self.transformer.weights.export_params()
        save_file({"transformer." + k: v.detach().float().cpu().contiguous() for k, v in P.items() if ".lora_" in k},
                  os.path.join(path, "pytorch_lora_weights.safetensors"))

    def state_dict(self, *args, **kwargs):
        """The LoongX layout train.py:214-217 saves and inference.py:46-52 loads: `transformer.<diffusers name>` with the
        peft-injected spellings (`.base_layer.weight`, `.lora_A.default.weight`) + the CS3 / DGF modules."""
        sd = super().state_dict(*args, **kwargs)
        P = self.transformer.weights.export_params()
        targets = {k.rsplit(".lora_", 1)[0] for k in P if ".lora_" in k}
        for k, v in P.items():
            stem, kind = k.rsplit(".", 1)
            if ".lora_" in k:
                mod, ab = k[:-len(".weight")].rsplit(".lora_", 1)
                sd[f"transformer.{mod}.lora_{ab}.default.weight"] = v
            elif stem in targets:
                sd[f"transformer.{stem}.base_layer.{kind}"] = v
            else:
                sd["transformer." + k] = v
        return sd

    def _load_transformer_state(self, tr) -> None:
        from loongx_b200.pipeline import init_lora_factors

        ranks = {v.shape[0] for k, v in tr.items() if k.endswith(".lora_A.weight")}
        if len(ranks) == 1:
            self.transformer.cfg.lora_rank = int(next(iter(ranks)))  # a checkpoint saved at another rank
        tr = {k: v.to(self.device) for k, v in tr.items()}
        init_lora_factors(tr, self.transformer.cfg, self.device)
        self.transformer.load_params(tr)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """LoongX `model.state_dict()` (inference.py:46-52): `transformer.*` keys (peft spellings accepted) replace the
        native DiT weights, everything else goes to the CS3 / DGF modules."""
        from loongx_b200.checkpoint import split_loongx_state_dict

        tr, rest = split_loongx_state_dict(state_dict)
        if tr:
            if self.staged:
                self._pending.append(("state", tr))
            else:
                self._load_transformer_state(tr)
        return super().load_state_dict(rest, strict=strict, assign=assign)

    def to_model_dtype(self, x: torch.Tensor) -> torch.Tensor:
        """fp32 conditioning output -> DiT dtype (native cast kernel)."""
        return cs3.cast_bf16(x) if x.dtype == torch.float32 else x

    def spatial_pyramid_pooling(self, x, output_size, adaptive=False):
        """model.py:479-511: zero-pad / truncate the last dim to `output_size`."""
        if adaptive:
            B, Cc, _ = x.shape
            out = torch.empty((B, Cc * output_size), device=x.device, dtype=torch.float32)
            cs3.adaptive_pool(cs3._f32(x), out, output_size, output_size, 1, 0)
            return out.view(B, Cc, output_size)
        return cs3.pad_truncate(x, output_size)

    def fuse_eeg(self, eeg_features, ppg_features):
        """model.py:731-755: fusion1 over the token axis of cat([eeg, DUAN(x=ppg, c=eeg)], dim=1)."""
        B, n_tok, Dm = eeg_features.shape
        cat = torch.empty((B, 2 * n_tok, Dm), device=eeg_features.device, dtype=torch.float32)
        cat[:, :n_tok].copy_(eeg_features)
        self.duan_norm1(ppg_features, eeg_features, out=cat[:, n_tok:])
        return cs3.token_axis_linear(self.fusion1[0], cat)

    def fuse_fnirs(self, fnirs_features, motion_features):
        """model.py:757-779: fusion2(cat([fnirs, DUAN_1(x=fnirs, c=motion)], -1))."""
        B, Dm = fnirs_features.shape
        cat = torch.empty((B, 2 * Dm), device=fnirs_features.device, dtype=torch.float32)
        cat[:, :Dm].copy_(fnirs_features)
        fused = self.duan_norm2(fnirs_features.unsqueeze(1), motion_features.unsqueeze(1))
        cat[:, Dm:].copy_(fused.squeeze(1))
        return cs3.gemv(self.fusion2[0].weight, self.fusion2[0].bias, cat)

    # ---- training step (model.py:513-729) -------------------------------------------------------------------------
    def _trainer(self, B, n_txt, n_img, n_cond):
        from loongx_b200.train import DitTrainer

        enc = self._wants_encoder_grads()
        key = (B, n_txt, n_img, n_cond, id(self.transformer.weights), enc)
        if getattr(self, "_trainer_key", None) != key:
            from loongx_b200 import cs3_bwd as CB

            self._trainer_obj = None  # free the old trainer's activations before building the new one
            extra = CB.grad_elements(CB.trainable_parameters(self)) if enc else 0
            self._trainer_obj = DitTrainer(self.transformer.weights, B, n_txt, n_img, n_cond, model_config=self.model_config,
                                           input_grads=enc, extra_grad_elems=extra)
            self._trainer_key = key
        return self._trainer_obj

    def _wants_encoder_grads(self) -> bool:
        """The reference's step() leaves gradients on every CS3 / DGF parameter (they are inside the autograd graph,
        model.py:656-701) and DDP all-reduces them with the LoRA factors; `self.encoder_grads = False` skips that half (the
        reference's optimizer never updates those parameters, model.py:541)."""
        return bool(self.use_brain_condition and getattr(self, "encoder_grads", True))

    def _micro_batch(self, B: int, S: int) -> int:
        """Samples per native forward/backward: the whole batch when its block activations fit in the free HBM (no
        recompute), else the largest divisor of B that does (the step then accumulates gradients over B / mb micro-
        batches, same result); B itself (with per-block recompute) when not even one sample fits.  `self.micro_batch`
        (int) overrides."""
        from loongx_b200.train import DitTrainer

        forced = getattr(self, "micro_batch", None)
        if forced:
            return int(forced)
        cache = self.__dict__.setdefault("_mb_cache", {})  # decided once per geometry (the trainer then owns that memory)
        if (B, S) not in cache:
            pad = lambda n: (n + 127) // 128 * 128  # noqa: E731
            w = self.transformer.weights
            cache[(B, S)] = B
            for mb in sorted((d for d in range(1, B + 1) if B % d == 0), reverse=True):
                if DitTrainer.activations_fit(w, mb, pad(S) + 256):
                    cache[(B, S)] = mb
                    break
        return cache[(B, S)]

    @property
    def last_t(self):
        """model.py:728 / callbacks.py:59: mean timestep of the last step() (None before the first one)."""
        v = self.__dict__.get("_last_t")
        return None if v is None else float(v)

    @property
    def lora_layers(self):
        """model.py:513-524: the LoRA factors (fp32 masters) — the only parameters the reference's optimizer trains.  One
        stable list of nn.Parameters per weight set: available before the first step() (Lightning calls
        configure_optimizers first, model.py:533) and shared by every trainer, whatever the batch geometry."""
        return self.transformer.lora_parameters()

    def configure_optimizers(self):
        """model.py:533-558 (Prodigy is a third-party optimiser that is not installed here)."""
        opt = self.optimizer_config
        self.trainable_params = self.lora_layers
        if opt["type"] == "AdamW":
            return torch.optim.AdamW(self.trainable_params, **opt["params"])
        if opt["type"] == "SGD":
            return torch.optim.SGD(self.trainable_params, **opt["params"])
        if opt["type"] == "Prodigy":
            try:
                import prodigyopt
            except ImportError as e:  # model.py:547-551 imports it at module level
                raise NotImplementedError("optimizer type 'Prodigy' needs the third-party `prodigyopt` package") from e
            return prodigyopt.Prodigy(self.trainable_params, **opt["params"])
        raise NotImplementedError(opt["type"])

    def _step_conditioning(self, prompt_embeds, pooled, eeg, fnirs, ppg, motion):
        """model.py:656-701: the *step* fuse order — DUAN(x = brain, c = text), cat, fusion3 / fusion4, residual."""
        pe_b = po_b = None
        if eeg is not None:
            e = self.eeg_projection(self.spatial_pyramid_pooling(cs3._f32(eeg), self.eeg_fixed_length))
            pe_b = self.fuse_eeg(e, self.ppg_projection(self.spatial_pyramid_pooling(cs3._f32(ppg), self.ppg_fixed_length))) \
                if ppg is not None else e
        if fnirs is not None:
            f = self.fnirs_projection(self.spatial_pyramid_pooling(cs3._f32(fnirs), self.fnirs_fixed_length))
            po_b = self.fuse_fnirs(f, self.motion_projection(self.spatial_pyramid_pooling(cs3._f32(motion), self.motion_fixed_length))) \
                if motion is not None else f
        if pe_b is None or po_b is None:
            raise ValueError("step(): use_brain_condition needs eeg and fnirs (model.py:680-701 reads both embeddings)")
        if not self.fuse_flag:  # model.py:699-701
            return self.to_model_dtype(pe_b), self.to_model_dtype(po_b)
        to32 = lambda x: cs3._f32(x) if x.dtype == torch.float32 else cs3.cast_f32(x.contiguous())  # noqa: E731
        pe32, po32 = to32(prompt_embeds), to32(pooled)
        B, n_tok, Dm = pe32.shape
        cat = torch.empty((B, 2 * n_tok, Dm), device=pe32.device, dtype=torch.float32)
        cat[:, :n_tok].copy_(pe32)
        self.duan_norm_prompt(pe_b, pe32, out=cat[:, n_tok:])
        pe = cs3.token_axis_linear(self.fusion3[0], cat, residual=pe32)
        fp = self.duan_norm_pooled(po_b.unsqueeze(1), po32.unsqueeze(1)).squeeze(1)
        catp = torch.empty((B, 2 * po32.shape[1], 1), device=po32.device, dtype=torch.float32)
        catp[:, :po32.shape[1], 0].copy_(po32)
        catp[:, po32.shape[1]:, 0].copy_(fp)
        po = cs3.token_axis_linear(self.fusion4[0], catp, residual=po32.unsqueeze(2).contiguous()).squeeze(2)
        return self.to_model_dtype(pe), self.to_model_dtype(po)

    def step(self, batch):
        """model.py:569-729 -> scalar loss whose `.backward()` runs the native backward (LoRA-factor gradients, mean
        all-reduced across ranks when torch.distributed is initialised).  `image` / `condition`: pictures (with a VAE
        attached, model.py:582,596) or latents [B,16,h,w]; text: `description` strings (with text encoders attached,
        model.py:585-587), or pre-computed `prompt_embeds` + `pooled_prompt_embeds` (or `description=(prompt_embeds,
        pooled)`).  `t` / `noise` may be supplied for reproducibility; otherwise they are drawn like model.py:590-591."""
        from src.flux.pipeline_tools import encode_images

        imgs, conditions = batch["image"], batch["condition"]
        position_delta = batch["position_delta"][0]
        position_scale = float(batch.get("position_scale", [1.0])[0])
        dev = self.device
        with torch.no_grad():
            x_0, img_ids = encode_images(self.flux_pipe, imgs)
            if "prompt_embeds" in batch:
                pe, po = batch["prompt_embeds"], batch["pooled_prompt_embeds"]
            elif isinstance(batch.get("description"), tuple):
                pe, po = batch["description"]
            elif self.flux_pipe.text_encoder is not None and self.flux_pipe.text_encoder_2 is not None:
                from src.flux.pipeline_tools import prepare_text_input

                prompts = batch["description"]  # list of strings, what SeedDataset / collate_step_batch produce
                pe, po, _ = prepare_text_input(self.flux_pipe, [prompts] if isinstance(prompts, str) else list(prompts))
            else:
                raise NotImplementedError("step(): `description` holds strings but no text encoders are attached "
                                          "(OminiModel(flux_pipe_id=<FLUX directory>) or flux_pipe.attach_text_encoders); "
                                          "alternatively put prompt_embeds / pooled_prompt_embeds in the batch")
            pe, po = pe.to(dev), po.to(dev)
            B = x_0.shape[0]
            t = batch["t"].to(dev).float() if "t" in batch else torch.sigmoid(torch.randn((B,), device=dev))
            x_1 = batch["noise"].to(dev).to(x_0.dtype) if "noise" in batch else torch.randn_like(x_0)
            condition_latents, condition_ids = encode_images(self.flux_pipe, conditions)
            condition_ids = condition_ids.float()
            condition_ids[:, 1] += position_delta[0]
            condition_ids[:, 2] += position_delta[1]
            if position_scale != 1.0:
                scale_bias = (position_scale - 1.0) / 2
                condition_ids[:, 1:] *= position_scale
                condition_ids[:, 1:] += scale_bias
            cond_ctx = None
            if self.use_brain_condition and self._wants_encoder_grads():
                from loongx_b200 import cs3_bwd as CB

                self._step_count = getattr(self, "_step_count", 0) + 1
                pe, po, cond_ctx = CB.step_conditioning_train(self, pe, po, batch.get("eeg"), batch.get("fnirs"),
                                                              batch.get("ppg"), batch.get("motion"), training=self.training,
                                                              seed=1000003 * int(batch.get("dropout_seed", 0)) + self._step_count)
            elif self.use_brain_condition:
                pe, po = self._step_conditioning(pe, po, batch.get("eeg"), batch.get("fnirs"), batch.get("ppg"),
                                                 batch.get("motion"))
            text_ids = torch.zeros(pe.shape[1], 3, device=dev)
        mb = self._micro_batch(B, pe.shape[1] + x_0.shape[1] + condition_latents.shape[1])
        if cond_ctx is not None:
            mb = min(mb, max(d for d in range(1, 9) if B % d == 0))  # input-gradient kernels take <= 8 samples per pass
        tr = self._trainer(mb, pe.shape[1], x_0.shape[1], condition_latents.shape[1])
        enc = None
        if cond_ctx is not None:
            from loongx_b200 import cs3_bwd as CB
            from loongx_b200.train import EncoderBackward

            enc = EncoderBackward(self, cond_ctx, CB.trainable_parameters(self), tr)
        if mb == B:
            loss = tr.step_loss(x_0.contiguous(), x_1.contiguous(), t, condition_latents, pe, po, text_ids, img_ids.float(),
                                condition_ids, 1.0, enc=enc)
        else:  # gradient accumulation over micro-batches whose activations fit in HBM without recompute
            chunks = [(x_0[i:i + mb].contiguous(), x_1[i:i + mb].contiguous(), t[i:i + mb], condition_latents[i:i + mb],
                       pe[i:i + mb], po[i:i + mb], text_ids, img_ids.float(), condition_ids, 1.0) for i in range(0, B, mb)]
            loss = tr.step_loss_micro(chunks, enc=enc)
        # model.py:728 reads `t.mean().item()` here; kept on the device and converted when `last_t` is read (the callback
        # does, once per log interval): a host sync at this point would leave the GPU idle while `loss.backward()` and the
        # optimizer are enqueued (12 ms of a 760 ms step at per-GPU batch 8)
        self._last_t = t.mean()
        return loss

    def training_step(self, batch, batch_idx=0):
        step_loss = self.step(batch)
        v = float(step_loss.detach())
        self.log_loss = v if not hasattr(self, "log_loss") else self.log_loss * 0.95 + v * 0.05
        return step_loss
