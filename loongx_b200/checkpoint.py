"""Checkpoint readers: the three on-disk formats the reference consumes, mapped onto the flat
diffusers-named parameter dict that `DitWeights` packs.

  * diffusers FLUX.1-dev directory (`FluxPipeline.from_pretrained(flux_path)`, model.py:398-400):
        <flux_path>/transformer/config.json
        <flux_path>/transformer/diffusion_pytorch_model.safetensors            (single file)   or
        <flux_path>/transformer/diffusion_pytorch_model.safetensors.index.json + ...-0000i-of-0000n.safetensors (shards)
  * peft LoRA file written by `FluxPipeline.save_lora_weights` (model.py:526-531): `pytorch_lora_weights.safetensors`
    with keys `transformer.<module>.lora_A.weight` / `.lora_B.weight`
  * LoongX full state dict (`torch.save(model.state_dict())`, inference.py:43-53): `transformer.<module>...` with the
    peft-injected spellings `.base_layer.weight` / `.lora_A.default.weight`, plus the CS3 / DGF modules
    (`eeg_projection.*`, `duan_norm_prompt.*`, `fusion1.*`, ...)

Host-side I/O and key renaming only; tensors are moved to the device one at a time (a 24 GB checkpoint never has to fit
in host memory twice) and the packing / LoRA merge happens in DitWeights.
"""
from __future__ import annotations

import json
import os
import re
from typing import Dict, Iterable, Tuple

import torch

from .config import FluxConfig, linear_shapes, lora_targets, rmsnorm_names

_CFG_KEYS = ("num_layers", "num_single_layers", "num_attention_heads", "attention_head_dim", "in_channels",
             "joint_attention_dim", "pooled_projection_dim", "guidance_embeds")


def config_from_diffusers(cfg_json: dict, lora_rank: int = 4, lora_alpha: float = 4.0) -> FluxConfig:
    kw = {k: cfg_json[k] for k in _CFG_KEYS if k in cfg_json}
    if "axes_dims_rope" in cfg_json:
        kw["axes_dims_rope"] = tuple(cfg_json["axes_dims_rope"])
    cfg = FluxConfig(lora_rank=lora_rank, lora_alpha=lora_alpha, **kw)
    cfg.validate()
    return cfg


def expected_keys(cfg: FluxConfig) -> set:
    keys = set(rmsnorm_names(cfg))
    for name in linear_shapes(cfg):
        keys.add(name + ".weight")
        keys.add(name + ".bias")
    return keys


def _safetensors_files(tdir: str) -> Iterable[str]:
    idx = os.path.join(tdir, "diffusion_pytorch_model.safetensors.index.json")
    if os.path.exists(idx):
        with open(idx) as f:
            files = sorted(set(json.load(f)["weight_map"].values()))
        return [os.path.join(tdir, x) for x in files]
    single = os.path.join(tdir, "diffusion_pytorch_model.safetensors")
    if os.path.exists(single):
        return [single]
    raise FileNotFoundError(f"no diffusion_pytorch_model*.safetensors under {tdir}")


def read_diffusers_transformer(flux_path: str, device="cpu", dtype=torch.bfloat16, lora_rank: int = 4,
                               lora_alpha: float = 4.0) -> Tuple[FluxConfig, Dict[str, torch.Tensor]]:
    """-> (FluxConfig, {diffusers parameter name: tensor on `device`}); validates names and shapes."""
    from safetensors import safe_open

    tdir = os.path.join(flux_path, "transformer") if os.path.isdir(os.path.join(flux_path, "transformer")) else flux_path
    with open(os.path.join(tdir, "config.json")) as f:
        cfg = config_from_diffusers(json.load(f), lora_rank, lora_alpha)
    P: Dict[str, torch.Tensor] = {}
    for fn in _safetensors_files(tdir):
        with safe_open(fn, framework="pt", device="cpu") as sf:
            for k in sf.keys():
                P[k] = sf.get_tensor(k).to(device=device, dtype=dtype)
    check_transformer_params(P, cfg)
    return cfg, P


def check_transformer_params(P: Dict[str, torch.Tensor], cfg: FluxConfig) -> None:
    want = expected_keys(cfg)
    have = {k for k in P if ".lora_" not in k}
    missing, unexpected = sorted(want - have), sorted(have - want)
    if missing or unexpected:
        raise KeyError(f"transformer checkpoint does not match the config: missing {missing[:4]} ({len(missing)}), "
                       f"unexpected {unexpected[:4]} ({len(unexpected)})")
    for name, (o, i) in linear_shapes(cfg).items():
        if tuple(P[name + ".weight"].shape) != (o, i) or tuple(P[name + ".bias"].shape) != (o,):
            raise ValueError(f"{name}: expected weight {(o, i)}, got {tuple(P[name + '.weight'].shape)}")


_PEFT = re.compile(r"^(?:transformer\.)?(.*)\.lora_([AB])(?:\.[A-Za-z0-9_]+)?\.weight$")


def read_peft_lora(path: str, device="cpu", dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """`save_lora_weights` output (directory or .safetensors file) -> {"<module>.lora_A.weight": [r, in], ...}."""
    from safetensors import safe_open

    fn = os.path.join(path, "pytorch_lora_weights.safetensors") if os.path.isdir(path) else path
    out: Dict[str, torch.Tensor] = {}
    with safe_open(fn, framework="pt", device="cpu") as sf:
        for k in sf.keys():
            m = _PEFT.match(k)
            if m is None:
                raise KeyError(f"unrecognised LoRA key {k!r}")
            out[f"{m.group(1)}.lora_{m.group(2)}.weight"] = sf.get_tensor(k).to(device=device, dtype=dtype)
    return out


def apply_lora(P: Dict[str, torch.Tensor], lora: Dict[str, torch.Tensor], cfg: FluxConfig) -> None:
    """Insert LoRA factors into the flat dict (in place); every factor must belong to a LoRA target of this build."""
    targets = set(lora_targets(cfg))
    shapes = linear_shapes(cfg)
    ranks = set()
    for k, v in lora.items():
        stem, ab = k.rsplit(".lora_", 1)
        if stem not in targets:
            raise KeyError(f"{stem} is not a LoRA target of train/config/seed_512.yaml:38")
        o, i = shapes[stem]
        r = v.shape[0] if ab.startswith("A") else v.shape[1]
        if (ab.startswith("A") and v.shape[1] != i) or (ab.startswith("B") and v.shape[0] != o):
            raise ValueError(f"{k}: shape {tuple(v.shape)} does not fit Linear({i}, {o})")
        ranks.add(r)
        P[k] = v
    if len(ranks) > 1:
        raise ValueError(f"mixed LoRA ranks {sorted(ranks)}")
    if ranks:
        cfg.lora_rank = ranks.pop()


def split_loongx_state_dict(sd: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """LoongX `model.state_dict()` -> (transformer params in diffusers naming incl. LoRA factors, everything else)."""
    if "state_dict" in sd and isinstance(sd["state_dict"], dict):  # inference.py:46-49
        sd = sd["state_dict"]
    tr, rest = {}, {}
    for k, v in sd.items():
        if k.startswith("flux_pipe."):
            continue  # aliases of transformer.* when Lightning registers the pipeline
        if not k.startswith("transformer."):
            rest[k] = v
            continue
        name = k[len("transformer."):]
        name = name.replace(".base_layer.", ".")
        name = re.sub(r"\.lora_([AB])\.[A-Za-z0-9_]+\.weight$", r".lora_\1.weight", name)
        tr[name] = v
    return tr, rest


def write_diffusers_transformer(path: str, cfg: FluxConfig, P: Dict[str, torch.Tensor], shards: int = 2) -> None:
    """Write a diffusers-format transformer directory (used by the tests and by users exporting synthetic weights)."""
    from safetensors.torch import save_file

    tdir = os.path.join(path, "transformer")
    os.makedirs(tdir, exist_ok=True)
    cj = {k: getattr(cfg, k) for k in _CFG_KEYS}
    cj.update(axes_dims_rope=list(cfg.axes_dims_rope), patch_size=1, _class_name="FluxTransformer2DModel")
    with open(os.path.join(tdir, "config.json"), "w") as f:
        json.dump(cj, f, indent=1)
    keys = sorted(k for k in P if ".lora_" not in k)
    if shards <= 1:
        save_file({k: P[k].contiguous().cpu() for k in keys}, os.path.join(tdir, "diffusion_pytorch_model.safetensors"))
        return
    per = (len(keys) + shards - 1) // shards
    weight_map = {}
    for s in range(shards):
        fn = f"diffusion_pytorch_model-{s + 1:05d}-of-{shards:05d}.safetensors"
        part = keys[s * per:(s + 1) * per]
        save_file({k: P[k].contiguous().cpu() for k in part}, os.path.join(tdir, fn))
        weight_map.update({k: fn for k in part})
    with open(os.path.join(tdir, "diffusion_pytorch_model.safetensors.index.json"), "w") as f:
        json.dump({"metadata": {}, "weight_map": weight_map}, f)
