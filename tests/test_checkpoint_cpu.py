"""CPU tests of the checkpoint readers: diffusers-format FLUX transformer directory (single file and
sharded + index), peft LoRA file (`save_lora_weights` layout) and the LoongX full state dict (peft-injected key
spellings, inference.py:43-53).  Bit-exact round trips; no GPU."""
import os

import pytest
import torch

from loongx_b200 import checkpoint as CK
from loongx_b200.config import FluxConfig, linear_shapes, lora_targets, rmsnorm_names

TINY = dict(num_layers=1, num_single_layers=2, num_attention_heads=2, joint_attention_dim=64, pooled_projection_dim=32)


def _params(cfg, lora=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, (o, i) in linear_shapes(cfg).items():
        P[name + ".weight"] = torch.randn(o, i, generator=g).bfloat16()
        P[name + ".bias"] = torch.randn(o, generator=g).bfloat16()
    for n in rmsnorm_names(cfg):
        P[n] = torch.randn(128, generator=g).bfloat16()
    if lora:
        for name in lora_targets(cfg):
            o, i = linear_shapes(cfg)[name]
            P[name + ".lora_A.weight"] = torch.randn(cfg.lora_rank, i, generator=g)
            P[name + ".lora_B.weight"] = torch.randn(o, cfg.lora_rank, generator=g)
    return P


@pytest.mark.parametrize("shards", [1, 3])
def test_diffusers_directory_round_trip(tmp_path, shards):
    cfg = FluxConfig(**TINY)
    P = _params(cfg)
    CK.write_diffusers_transformer(str(tmp_path), cfg, P, shards=shards)
    names = os.listdir(tmp_path / "transformer")
    assert ("diffusion_pytorch_model.safetensors.index.json" in names) == (shards > 1)
    cfg2, P2 = CK.read_diffusers_transformer(str(tmp_path), lora_rank=8, lora_alpha=16.0)
    assert (cfg2.num_layers, cfg2.num_single_layers, cfg2.inner_dim, cfg2.joint_attention_dim) == (1, 2, 256, 64)
    assert cfg2.lora_rank == 8 and cfg2.lora_alpha == 16.0 and cfg2.axes_dims_rope == (16, 56, 56)
    assert set(P2) == set(P) == CK.expected_keys(cfg)
    assert all(torch.equal(P[k], P2[k]) for k in P)


def test_diffusers_directory_rejects_mismatch(tmp_path):
    cfg = FluxConfig(**TINY)
    P = _params(cfg)
    P.pop("proj_out.bias")
    CK.write_diffusers_transformer(str(tmp_path), cfg, P, shards=1)
    with pytest.raises(KeyError):
        CK.read_diffusers_transformer(str(tmp_path))
    with pytest.raises(FileNotFoundError):
        CK.read_diffusers_transformer(str(tmp_path / "nope"))


def test_peft_lora_file_and_apply(tmp_path):
    from safetensors.torch import save_file

    cfg = FluxConfig(**TINY, lora_rank=4)
    P = _params(cfg, lora=True)
    lora = {k: v for k, v in P.items() if ".lora_" in k}
    save_file({"transformer." + k: v for k, v in lora.items()}, str(tmp_path / "pytorch_lora_weights.safetensors"))
    got = CK.read_peft_lora(str(tmp_path))
    assert set(got) == set(lora) and all(torch.equal(got[k], lora[k]) for k in lora)
    base = {k: v for k, v in P.items() if ".lora_" not in k}
    cfg.lora_rank = 0
    CK.apply_lora(base, got, cfg)
    assert cfg.lora_rank == 4 and set(base) == set(P)
    with pytest.raises(KeyError):  # context-stream Linear is not a LoRA target (seed_512.yaml:38)
        CK.apply_lora(dict(base), {"transformer_blocks.0.attn.add_q_proj.lora_A.weight": torch.zeros(4, 256)}, cfg)
    with pytest.raises(ValueError):
        CK.apply_lora(dict(base), {"x_embedder.lora_A.weight": torch.zeros(4, 7)}, cfg)


def test_loongx_state_dict_split_handles_peft_spellings():
    cfg = FluxConfig(**TINY, lora_rank=4)
    P = _params(cfg, lora=True)
    targets = set(lora_targets(cfg))
    sd = {}
    for k, v in P.items():
        stem, kind = k.rsplit(".", 1)
        if ".lora_" in k:
            mod, ab = k[:-len(".weight")].rsplit(".lora_", 1)
            sd[f"transformer.{mod}.lora_{ab}.default.weight"] = v
        elif stem in targets:
            sd[f"transformer.{stem}.base_layer.{kind}"] = v
        else:
            sd["transformer." + k] = v
    sd["eeg_projection.projection.1.weight"] = torch.zeros(3)
    sd["duan_norm_prompt.gate.0.bias"] = torch.ones(2)
    sd["flux_pipe.transformer.proj_out.bias"] = torch.ones(1)
    tr, rest = CK.split_loongx_state_dict({"state_dict": sd})
    assert set(tr) == set(P) and all(torch.equal(tr[k], P[k]) for k in P)
    assert set(rest) == {"eeg_projection.projection.1.weight", "duan_norm_prompt.gate.0.bias"}
    CK.check_transformer_params(tr, cfg)
