"""Drop-in for the reference's src/train/model.py: OminiModel + the CS3 / DGF modules (model.py:16-1035).

Constructor signature and the attribute surface consumed by generate() (model.py:377-462) are kept;
the sub-modules are the native-kernel shells from loongx_b200.cs3.  The Lightning training harness, optimisers and
checkpoint serialisation (model.py:513-567, 780-943) are outside this build's scope.
"""
from typing import Optional

import torch
import torch.nn as nn

from loongx_b200 import cs3
from loongx_b200.config import FluxConfig
from loongx_b200.cs3 import DUAN, EEGEncoder, FeaturePyramidPooling, FNIRSEncoder, MotionEncoder, PPGEncoder  # noqa: F401
from loongx_b200.pipeline import NativeFluxPipeline, NativeFluxTransformer


class OminiModel(nn.Module):
    def __init__(self, flux_pipe_id, lora_path: str = None, lora_config: dict = None, device: str = "cuda",
                 dtype: torch.dtype = torch.bfloat16, model_config: dict = {}, optimizer_config: dict = None,
                 gradient_checkpointing: bool = False, use_brain_condition: bool = True, fuse_flag: bool = True,
                 seed: int = 1234):
        """`flux_pipe_id`: a FluxConfig (random-init weights of that architecture, seeded) or the string "synthetic"
        (FLUX.1-dev geometry).  Loading real diffusers / peft checkpoints is SURVEY.md §8f.1 (next)."""
        super().__init__()
        if dtype != torch.bfloat16:
            raise NotImplementedError("the native DiT computes in bf16 (fp32 accumulate); CS3/DGF run in float32")
        if isinstance(flux_pipe_id, FluxConfig):
            cfg = flux_pipe_id
        elif flux_pipe_id == "synthetic":
            cfg = FluxConfig()
        else:
            raise NotImplementedError(f"checkpoint loading ({flux_pipe_id!r}) is not built yet: pass a FluxConfig or 'synthetic'")
        if lora_config is not None:
            cfg.lora_rank = int(lora_config.get("r", cfg.lora_rank))
            cfg.lora_alpha = float(lora_config.get("lora_alpha", cfg.lora_alpha))
        if lora_path:
            raise NotImplementedError  # model.py:517 raises as well
        self.model_config = model_config
        self.optimizer_config = optimizer_config
        self._dtype = dtype
        self._device = torch.device(device)
        torch.manual_seed(seed)
        self.transformer = NativeFluxTransformer(cfg, device=device, seed=seed)
        self.transformer.gradient_checkpointing = gradient_checkpointing
        self.flux_pipe = NativeFluxPipeline(self.transformer)
        self.fuse_flag = fuse_flag
        self.use_brain_condition = use_brain_condition
        self.eeg_fixed_length, self.fnirs_fixed_length, self.ppg_fixed_length, self.motion_fixed_length = 4096, 512, 256, 128

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

    @property
    def device(self):
        return self._device

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

    def step(self, batch):
        raise NotImplementedError("the training step (model.py:569-729: backward kernels + DDP all-reduce) is not built "
                                  "in this round; see DESIGN.md 'What comes next'")

    training_step = step
