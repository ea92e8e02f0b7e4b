"""GPU parity of the reference-granularity module surface `src.flux.block.{attn_forward, block_forward,
single_block_forward}` (block.py:7-339) against the oracle restatement, incl. ragged token counts, the
mask variants, c_factor and latent_lora.  Tolerance as in tests/test_dit_gpu.py (bf16 storage): relL2 <= 1.5e-2."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-30)).item()


def _setup(nt=128, h=8, w=16, B=2, seed=0):
    from oracle import flux_dit as O
    from loongx_b200.config import FluxConfig
    from loongx_b200.pipeline import NativeFluxTransformer

    kw = dict(num_layers=2, num_single_layers=2, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
    ocfg, cfg = O.FluxConfig(**kw), FluxConfig(**kw)
    P = O.init_params(ocfg, seed=1234, dtype=torch.float32, device="cpu", w_std=0.05, bias_std=0.05, lora_b_std=0.05)
    Pb = {k: v.to(torch.bfloat16).to(DEV) for k, v in P.items()}
    P32 = {k: v.float() for k, v in Pb.items()}
    tr = NativeFluxTransformer(cfg, params=dict(Pb), device=DEV)
    g = torch.Generator().manual_seed(seed)
    ni = h * w
    D = ocfg.inner_dim
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).bfloat16().to(DEV)  # noqa: E731
    ids = torch.zeros(h, w, 3)
    ids[..., 1] += torch.arange(h)[:, None]
    ids[..., 2] += torch.arange(w)[None, :]
    ids = ids.reshape(-1, 3).to(DEV)
    cids = ids.clone()
    cids[:, 2] -= w
    x = dict(h=r(B, ni, D), e=r(B, nt, D), c=r(B, ni, D), temb=r(B, D), ctemb=r(B, D),
             rope=O.rope_tables(torch.cat([torch.zeros(nt, 3, device=DEV), ids], 0)), crope=O.rope_tables(cids))
    return O, ocfg, P32, tr, x


@pytest.mark.parametrize("nt,hw,mc,cf", [(128, (8, 16), {}, None), (77, (6, 10), {}, None), (128, (8, 16), {"latent_lora": True}, None),
                                         (128, (8, 16), {"union_cond_attn": False}, None), (100, (8, 16), {}, 1.6)])
def test_block_and_single_block_forward(nt, hw, mc, cf):
    from src.flux.block import block_forward, single_block_forward

    O, ocfg, P32, tr, x = _setup(nt, *hw)
    f = lambda t: t.float()  # noqa: E731
    if cf is not None:
        for blk in tr.transformer_blocks + tr.single_transformer_blocks:
            blk.attn.c_factor = torch.ones(1, 1) * cf
    with torch.no_grad():
        e_ref, h_ref, c_ref = O.block_forward(P32, ocfg, 1, f(x["h"]), f(x["e"]), f(x["c"]), f(x["temb"]), f(x["ctemb"]),
                                              x["crope"], x["rope"], mc, cf)
    e, h, c = block_forward(tr.transformer_blocks[1], x["h"], x["e"], x["c"], x["temb"], x["ctemb"], cond_rotary_emb=x["crope"],
                            image_rotary_emb=x["rope"], model_config=mc)
    errs = (_rel(e, e_ref), _rel(h, h_ref), _rel(c, c_ref))
    assert e.shape == x["e"].shape and h.shape == x["h"].shape and c.dtype == x["c"].dtype
    # no condition
    with torch.no_grad():
        e_ref2, h_ref2, _ = O.block_forward(P32, ocfg, 0, f(x["h"]), f(x["e"]), None, f(x["temb"]), None, None, x["rope"], mc, None)
    e2, h2, c2 = block_forward(tr.transformer_blocks[0], x["h"], x["e"], None, x["temb"], None, image_rotary_emb=x["rope"],
                               model_config=mc)
    assert c2 is None
    # single block on cat([txt, img])
    xs = torch.cat([x["e"], x["h"]], 1)
    with torch.no_grad():
        s_ref, sc_ref = O.single_block_forward(P32, ocfg, 1, f(xs), f(x["temb"]), x["rope"], f(x["c"]), f(x["ctemb"]), x["crope"],
                                               mc, cf)
        s_ref2 = O.single_block_forward(P32, ocfg, 0, f(xs), f(x["temb"]), x["rope"], None, None, None, mc, None)
    s, sc = single_block_forward(tr.single_transformer_blocks[1], xs, x["temb"], image_rotary_emb=x["rope"],
                                 condition_latents=x["c"], cond_temb=x["ctemb"], cond_rotary_emb=x["crope"], model_config=mc)
    s2 = single_block_forward(tr.single_transformer_blocks[0], xs, x["temb"], image_rotary_emb=x["rope"], model_config=mc)
    errs += (_rel(e2, e_ref2), _rel(h2, h_ref2), _rel(s, s_ref), _rel(sc, sc_ref), _rel(s2, s_ref2))
    print(f"\n[block api nt={nt} hw={hw} {mc} cf={cf}] relL2 " + " ".join(f"{v:.3g}" for v in errs))
    assert max(errs) < 1.5e-2, errs


@pytest.mark.parametrize("nt,hw,mc", [(128, (8, 16), {}), (50, (6, 10), {"independent_condition": True})])
def test_attn_forward(nt, hw, mc):
    from src.flux.block import attn_forward

    O, ocfg, P32, tr, x = _setup(nt, *hw)
    f = lambda t: t.float()  # noqa: E731
    with torch.no_grad():
        h_ref, e_ref, c_ref = O.attn_forward(P32, ocfg, "transformer_blocks.0.attn", f(x["h"]), f(x["e"]), f(x["c"]), x["rope"],
                                             x["crope"], mc, None)
        xs = torch.cat([x["e"], x["h"]], 1)
        s_ref, sc_ref = O.attn_forward(P32, ocfg, "single_transformer_blocks.1.attn", f(xs), None, f(x["c"]), x["rope"], x["crope"],
                                       mc, None)
        s_ref2 = O.attn_forward(P32, ocfg, "single_transformer_blocks.0.attn", f(xs), None, None, x["rope"], None, mc, None)
    h, e, c = attn_forward(tr.transformer_blocks[0].attn, x["h"], x["e"], x["c"], image_rotary_emb=x["rope"],
                           cond_rotary_emb=x["crope"], model_config=mc)
    s, sc = attn_forward(tr.single_transformer_blocks[1].attn, xs, condition_latents=x["c"], image_rotary_emb=x["rope"],
                         cond_rotary_emb=x["crope"], model_config=mc)
    s2 = attn_forward(tr.single_transformer_blocks[0].attn, xs, image_rotary_emb=x["rope"], model_config=mc)
    errs = (_rel(h, h_ref), _rel(e, e_ref), _rel(c, c_ref), _rel(s, s_ref), _rel(sc, sc_ref), _rel(s2, s_ref2))
    print(f"\n[attn_forward nt={nt} hw={hw} {mc}] relL2 " + " ".join(f"{v:.3g}" for v in errs))
    assert max(errs) < 1.5e-2, errs
    with pytest.raises(TypeError):
        attn_forward(tr.single_transformer_blocks[0].attn, xs, encoder_hidden_states=x["e"])
