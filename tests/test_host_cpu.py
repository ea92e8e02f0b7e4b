"""CPU tests of the host logic and of the C-ABI surface (library loads, exports every declared symbol, argument
validation fails loudly) — no compute call needs a GPU here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge

    ge.build()
    from loongx_b200 import _lib as L

    syms = L.exported_symbols()
    assert len(syms) >= 30 and "lx_dit_step" in syms and "lx_duan_forward" in syms
    for s in syms:
        assert hasattr(L.lib, s), f"{s} declared in include/loongx_b200.h but not exported"
    assert L.lib.lx_version() >= 100


def test_struct_layouts_match_the_header(tmp_path):
    from loongx_b200 import _lib as L
    from loongx_b200 import cs3, dit, text, vae

    pairs = {"lx_vae_im2col_desc_t": vae.Im2colDesc, "lx_small_attn_desc_t": text.SmallAttnDesc,
             "lx_tile_meta_t": L.TileMeta, "lx_gemm_group_t": L.GemmGroup, "lx_gemm_segment_t": L.GemmSegment,
             "lx_gemm_desc_t": L.GemmDesc, "lx_attn_desc_t": L.AttnDesc, "lx_attn_bwd_desc_t": L.AttnBwdDesc, "lx_linear_t": dit.LxLinear,
             "lx_double_block_t": dit.LxDoubleBlock, "lx_single_block_t": dit.LxSingleBlock,
             "lx_dit_model_t": dit.LxDitModel, "lx_dit_plan_t": dit.LxDitPlan, "lx_sgemm_desc_t": cs3.SgemmDesc,
             "lx_duan_weights_t": cs3.DuanWeights}
    src = tmp_path / "sz.c"
    body = "\n".join(f'  printf("{n} %zu\\n", sizeof({n}));' for n in pairs)
    src.write_text(f'#include <stdio.h>\n#include "loongx_b200.h"\nint main(void) {{\n{body}\n  return 0;\n}}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        name, size = line.split()
        assert C.sizeof(pairs[name]) == int(size), f"{name}: ctypes {C.sizeof(pairs[name])} != C {size}"


def test_argument_validation_fails_loudly_without_gpu():
    from loongx_b200 import _lib as L

    assert L.lib.lx_gemm_bf16(None, None) == -1 and b"null descriptor" in L.lib.lx_last_error()
    d = L.GemmDesc()
    d.M, d.N, d.n_groups = 128, 250, 1  # N not a multiple of 8
    d.A = 16
    assert L.lib.lx_gemm_bf16(C.byref(d), None) == -1
    a = L.AttnDesc()
    a.q = a.k = a.v = a.out = a.out_row_base = 16
    a.B, a.H, a.S = 1, 1, 100  # S not a multiple of 128
    assert L.lib.lx_attention(C.byref(a), None) == -1 and b"multiple of 128" in L.lib.lx_last_error()
    with pytest.raises(L.LoongXNativeError):
        L.check(-1, "x")


def test_product_path_has_no_cpu_fallback():
    from loongx_b200 import _lib as L
    from loongx_b200 import cs3, ops

    with pytest.raises(L.LoongXNativeError):
        cs3.pad_truncate(torch.zeros(1, 4, 100), 128)
    with pytest.raises(AssertionError):
        ops.gemm(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(256, 64, dtype=torch.bfloat16), None,
                 torch.zeros(128, 256, dtype=torch.bfloat16))
    # the product never imports the oracle
    for dirpath, _, files in list(os.walk(os.path.join(ROOT, "loongx_b200"))) + list(os.walk(os.path.join(ROOT, "src"))):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f"{f} imports the oracle"


def test_tile_meta_and_row_base_are_inverse():
    from loongx_b200 import ops

    B, nt, ni, nc = 3, 256, 512, 384
    S = nt + ni + nc
    meta = ops.make_tile_meta(B, nt, ni, nc, "cpu")
    orb = ops.make_out_row_base(B, nt, ni, nc, "cpu")
    assert meta.shape == (B * S // 128, 4) and orb.shape == (B * S // 128,)
    seen = set()
    for t, (stream, b, seq_row, _) in enumerate(meta.tolist()):
        assert seq_row // S == b
        s0 = seq_row % S
        assert stream == (0 if s0 < nt else 1 if s0 < nt + ni else 2)
        assert orb[seq_row // 128].item() == t * 128  # (batch, sequence tile) -> row of that tile
        seen.add(seq_row)
    assert len(seen) == B * S // 128
    with pytest.raises(AssertionError):
        ops.make_tile_meta(1, 100, 128, 128, "cpu")


def test_scheduler_matches_closed_form_and_reference_sharding():
    from loongx_b200 import sampler as S

    sch = S.FlowMatchEulerDiscreteScheduler()
    sig = np.linspace(1.0, 1 / 28, 28)
    mu = S.calculate_shift(1024, 256, 4096, 0.5, 1.15)
    assert abs(mu - 0.63) < 1e-9
    ts, n = S.retrieve_timesteps(sch, 28, None, None, sig, mu=mu)
    assert n == 28 and ts.dtype == np.float32 and ts[0] == 1000.0 and sch.sigmas[-1] == 0.0
    ref = np.exp(mu) / (np.exp(mu) + (1 / sig - 1))
    assert np.allclose(sch.sigmas[:-1], ref, atol=1e-6)
    dts = [sch.advance() for _ in range(28)]
    assert abs(sum(dts) + 1.0) < 1e-5 and all(d < 0 for d in dts) and sch.step_index == 28
    with pytest.raises(ValueError):
        S.retrieve_timesteps(sch, 28, None, [1, 2], None, mu=mu)
    with pytest.raises(ValueError):
        sch.set_timesteps(28)  # dynamic shifting needs mu
    ids = S.latent_image_ids(4, 6)
    assert ids.shape == (24, 3) and ids[7].tolist() == [0.0, 1.0, 1.0]
    # inference.py:126-128: contiguous chunks, last rank takes the remainder
    for n_items, world in [(10, 4), (8, 8), (3, 2), (17, 8)]:
        got = [S.shard_range(n_items, r, world) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == n_items
        assert all(got[i][1] == got[i + 1][0] for i in range(world - 1))
        assert all(e - s == n_items // world for s, e in got[:-1])


def test_weight_packing_merges_lora_for_the_condition_rows():
    from loongx_b200.config import FluxConfig, linear_shapes, lora_targets
    from loongx_b200.dit import DitWeights, mask_mode_from_config, pack_linear, random_params

    cfg = FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
    P = random_params(cfg, "cpu", seed=5, lora_b_std=0.1)
    assert set(k[:-len(".lora_A.weight")] for k in P if k.endswith(".lora_A.weight")) == set(lora_targets(cfg))
    names = ["single_transformer_blocks.0.attn.to_q", "single_transformer_blocks.0.attn.to_k",
             "single_transformer_blocks.0.attn.to_v", "single_transformer_blocks.0.proj_mlp"]
    pl = pack_linear(P, names, cfg, "cpu")
    D = cfg.inner_dim
    assert pl.w.shape == (7 * D, D) and pl.w_lora.shape == (7 * D, D) and pl.bias.dtype == torch.float32
    a, b = P[names[1] + ".lora_A.weight"].float(), P[names[1] + ".lora_B.weight"].float()
    ref = (P[names[1] + ".weight"].float() + b @ a * (cfg.lora_alpha / cfg.lora_rank)).to(torch.bfloat16)
    assert torch.equal(pl.w_lora[D:2 * D], ref) and torch.equal(pl.w[D:2 * D], P[names[1] + ".weight"])
    ctx = pack_linear(P, ["transformer_blocks.0.attn.add_q_proj"], cfg, "cpu")
    assert ctx.w_lora is None  # text stream: no LoRA
    W = DitWeights(P, cfg, "cpu")
    assert W.model.mod_img.n == 6 * D and W.model.mod_single.n == 3 * D and W.sgl[0].qkv_mlp.n == 7 * D
    assert W.dbl[0].ff_up.w_lora is None and W.dbl[0].ff_down.w_lora is not None
    assert sum(o * i for o, i in linear_shapes(FluxConfig()).values()) > 11.8e9  # FLUX.1-dev: ~11.9 B parameters
    assert mask_mode_from_config({}) == 0 and mask_mode_from_config({"union_cond_attn": False, "independent_condition": True}) == 1
    assert mask_mode_from_config({"independent_condition": True}) == 2
    with pytest.raises(ValueError):
        FluxConfig(num_attention_heads=3).validate()


def test_training_kernels_validate_arguments_without_gpu():
    """The training-step entry points fail loudly on bad arguments before touching the device."""
    from loongx_b200 import _lib as L
    import loongx_b200.train  # noqa: F401  (argtypes)

    lib = L.lib
    assert lib.lx_gelu_fwd(16, 8, 16, 8, 4, 12, None) == -1          # cols not a multiple of 8
    assert lib.lx_lora_grad(16, 8, 16, 8, 16, 16, 16, 16, 4, 64, 64, 32, 1.0, 16, None) == -1 and b"rank" in lib.lx_last_error()
    assert lib.lx_flow_mse_loss(16, 16, 16, 16, None, 12, 1.0, None) == -1  # n not a multiple of 8
    b = L.AttnBwdDesc()
    b.q = b.k = b.v = b.d_out = b.lse = b.delta = b.dq = b.dk = b.dv = 16
    b.B, b.H, b.S, b.n_cond = 1, 1, 256, 100
    assert lib.lx_attention_bwd(C.byref(b), None) == -1 and b"n_cond" in lib.lx_last_error()
    a = L.AttnDesc()
    a.q = a.k = a.v = a.out = a.out_row_base = 16
    a.B, a.H, a.S = 1, 1, 256
    a.pad[0] = 130  # padding must be < 128
    assert lib.lx_attention(C.byref(a), None) == -1 and b"padding" in lib.lx_last_error()


def test_plan_padding_arithmetic_and_micro_batch_choice(monkeypatch):
    """Host-side geometry: streams padded to 128-token tiles; the training micro-batch is the largest divisor of the batch
    whose block activations fit in the free HBM."""
    from loongx_b200.config import FluxConfig
    from loongx_b200.dit import _pad128
    from loongx_b200.train import DitTrainer

    assert [_pad128(n) for n in (1, 77, 128, 129, 1024, 1600)] == [128, 128, 128, 256, 1024, 1664]
    cfg = FluxConfig()
    per_sample = DitTrainer.activation_bytes(cfg, 1, 2560)
    assert 15e9 < per_sample < 16.5e9 and DitTrainer.activation_bytes(cfg, 4, 2560) == 4 * per_sample

    class _W:  # the attributes activations_fit() reads
        device, named = "cuda", {}

        @staticmethod
        def param_bytes():
            return 55 * 10**9

    _W.cfg = cfg

    for free_gb, expect in ((180, [True, True, False]), (100, [True, False, False]), (60, [False, False, False])):
        monkeypatch.setattr(torch.cuda, "mem_get_info", lambda dev=None, f=free_gb: (f * 10**9, 192 * 10**9))
        assert [DitTrainer.activations_fit(_W, B, 2560) for B in (1, 4, 8)] == expect, free_gb


def test_checkpoint_key_renaming_is_total():
    """Every parameter name of the FLUX.1-dev geometry survives the LoongX <-> diffusers renaming both ways."""
    from loongx_b200 import checkpoint as CK
    from loongx_b200.config import FluxConfig, lora_targets

    cfg = FluxConfig()
    keys = sorted(CK.expected_keys(cfg))
    assert len(keys) == 2 * len(__import__("loongx_b200.config", fromlist=["linear_shapes"]).linear_shapes(cfg)) + 19 * 4 + 38 * 2
    targets = set(lora_targets(cfg))
    assert len(targets) == 1 + 19 * 6 + 38 * 6
    sd = {}
    for k in keys:
        stem, kind = k.rsplit(".", 1)
        sd[f"transformer.{stem}.base_layer.{kind}" if stem in targets else "transformer." + k] = k
    for t in targets:
        sd[f"transformer.{t}.lora_A.default.weight"] = t + ".lora_A.weight"
        sd[f"transformer.{t}.lora_B.default.weight"] = t + ".lora_B.weight"
    tr, rest = CK.split_loongx_state_dict(sd)
    assert not rest and all(k == v for k, v in tr.items()) and len(tr) == len(keys) + 2 * len(targets)


def test_integration_doc_stub_matches_the_header():
    """The ctypes stub INTEGRATION.md shows a reference maintainer mirrors lx_attn_desc_t field for field."""
    import ctypes as C
    from pathlib import Path

    from loongx_b200 import _lib as L

    doc = (Path(__file__).resolve().parent.parent / "INTEGRATION.md").read_text()
    code = doc[doc.index("class AttnDesc(C.Structure)"):doc.index("lib.lx_attention.argtypes")]
    ns = {"C": C}
    exec(code, ns)
    stub = ns["AttnDesc"]
    assert C.sizeof(stub) == C.sizeof(L.AttnDesc)
    assert [(n, C.sizeof(t)) for n, t in stub._fields_] == [(n, C.sizeof(t)) for n, t in L.AttnDesc._fields_]


def test_lora_context_managers_compose_with_the_per_forward_scale():
    """src/flux/lora_controller.py on a stand-in transformer (no GPU): enable_lora(activated=False) and set_lora_scale put a
    multiplier NEXT to the per-forward joint_attention_kwargs scale (peft's scale_lora_layers multiplies, lora_controller.py
    :5-75), restore it on exit (also on an exception), nest, and skip objects that are not native handles."""
    from src.flux.lora_controller import enable_lora, set_lora_scale

    class W:
        pass

    class Tr:
        def __init__(self):
            self.weights, self.log = W(), []

        def set_lora_outer(self, f):
            self.weights.lora_outer = float(f)
            self.log.append(float(f))

    class Block:  # what transformer.transformer_blocks[i] looks like to a caller
        def __init__(self, tr):
            self.transformer = tr

    tr = Tr()
    with enable_lora([Block(tr), Block(tr), object()], False):  # one weight set behind both handles: one switch
        assert tr.weights.lora_outer == 0.0
        with set_lora_scale([tr], 0.5):
            assert tr.weights.lora_outer == 0.0  # 0 * 0.5
        assert tr.weights.lora_outer == 0.0
    assert tr.weights.lora_outer == 1.0 and tr.log == [0.0, 0.0, 0.0, 1.0]
    with enable_lora([tr], True):  # activated: nothing changes
        assert tr.weights.lora_outer == 1.0
    with set_lora_scale([tr], 0.5):
        with set_lora_scale([Block(tr)], 0.5):
            assert tr.weights.lora_outer == 0.25
        assert tr.weights.lora_outer == 0.5
    assert tr.weights.lora_outer == 1.0
    try:
        with enable_lora([tr], False):
            raise KeyError("x")
    except KeyError:
        pass
    assert tr.weights.lora_outer == 1.0
