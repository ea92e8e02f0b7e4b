"""GPU parity of the native text encoders against oracle/text_encoders.py, which tests/test_text_cpu.py
pins against transformers' T5EncoderModel / CLIPTextModel.

Tolerances: bf16 weights, activations and residual stream with fp32 accumulation vs the fp32 oracle run on the same
bf16-rounded weights: kernel-level checks <= 1 bf16 ulp-ish (2e-2 abs on O(1) values), whole encoders relL2 <= 2e-2."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _round(P):
    return {k: v.to(torch.bfloat16).float() for k, v in P.items()}


@pytest.mark.parametrize("scalar", [0, 1])
def test_small_attention_kernel(scalar):
    """Both forms of lx_attention_small (mma.sync tiles = the default, scalar FMA = the cross-check) against torch."""
    from loongx_b200.text import _lib as lib

    lib.lx_debug_small_attention_scalar(scalar)
    try:
        _check_small_attention()
    finally:
        lib.lx_debug_small_attention_scalar(0)


def _check_small_attention():
    import ctypes as C

    from loongx_b200 import _lib as L
    from loongx_b200.text import SmallAttnDesc, _lib, _stream

    g = torch.Generator(device="cuda").manual_seed(0)
    for (B, H, S, causal, with_bias, scale) in ((2, 3, 77, True, False, 0.125), (1, 4, 200, False, True, 1.0),
                                                (1, 2, 512, False, True, 1.0), (2, 2, 33, True, True, 0.5),
                                                (3, 30, 50, True, False, 0.125), (1, 64, 96, False, True, 1.0)):
        inner = H * 64
        qkv = (torch.randn(B * S, 3 * inner, generator=g, device="cuda") * (0.4 if scale == 1.0 else 1.0)).to(torch.bfloat16)
        bias = torch.randn(H, S, S, generator=g, device="cuda") if with_bias else None
        out = torch.full((B * S, inner), float("nan"), dtype=torch.bfloat16, device="cuda")
        d = SmallAttnDesc()
        d.q, d.k, d.v = qkv.data_ptr(), qkv[:, inner:].data_ptr(), qkv[:, 2 * inner:].data_ptr()
        d.ldq = d.ldk = d.ldv = qkv.stride(0)
        d.out, d.ldo, d.bias = out.data_ptr(), out.stride(0), None if bias is None else bias.data_ptr()
        d.B, d.H, d.S, d.head_dim, d.causal, d.scale = B, H, S, 64, int(causal), scale
        if with_bias and S % 2 == 0:  # the distance-table form of the bias: [H, 2S-1], entry key - query + S - 1
            table = torch.randn(H, 2 * S - 1, generator=g, device="cuda")
            idx = torch.arange(S, device="cuda")[None, :] - torch.arange(S, device="cuda")[:, None] + S - 1
            bias = table[:, idx].contiguous()  # the equivalent full [H, S, S] bias for the torch reference
            d.bias, d.bias_relative = table.data_ptr(), 1
        L.check(_lib.lx_attention_small(C.byref(d), _stream()), "lx_attention_small")
        q, k, v = (t.float().view(B, S, H, 64).transpose(1, 2) for t in qkv.split(inner, dim=1))
        logits = q @ k.transpose(2, 3) * scale
        if bias is not None:
            logits = logits + bias[None]
        if causal:
            logits = logits + torch.full((S, S), float("-inf"), device="cuda").triu(1)
        want = (torch.softmax(logits, -1) @ v).transpose(1, 2).reshape(B * S, inner)
        assert torch.isfinite(out.float()).all()
        assert (out.float() - want).abs().max().item() < 2e-2, (B, H, S, causal)


def test_norm_embed_mul_kernels():
    from loongx_b200 import _lib as L
    from loongx_b200.text import _lib, _stream

    g = torch.Generator(device="cuda").manual_seed(1)
    for D in (128, 768, 4096):
        x = (torch.randn(37, D, generator=g, device="cuda") * 3 + 1).to(torch.bfloat16)
        gamma, beta = torch.randn(D, generator=g, device="cuda"), torch.randn(D, generator=g, device="cuda")
        out = torch.empty_like(x)
        L.check(_lib.lx_norm_rows(x.data_ptr(), D, gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), D, 37, D, 1e-5, 0, _stream()))
        want = torch.nn.functional.layer_norm(x.float(), (D,), gamma, beta, 1e-5)
        assert (out.float() - want).abs().max().item() < 4e-2
        L.check(_lib.lx_norm_rows(x.data_ptr(), D, gamma.data_ptr(), None, out.data_ptr(), D, 37, D, 1e-6, 1, _stream()))
        want = gamma * (x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-6))
        assert (out.float() - want).abs().max().item() < 4e-2
    table = torch.randn(50, 128, generator=g, device="cuda").to(torch.bfloat16)
    pos = torch.randn(7, 128, generator=g, device="cuda").to(torch.bfloat16)
    ids = torch.randint(0, 50, (3, 7), generator=g, device="cuda").to(torch.int32)
    out = torch.empty(21, 128, dtype=torch.bfloat16, device="cuda")
    L.check(_lib.lx_embed_rows(table.data_ptr(), ids.data_ptr(), None, 0, out.data_ptr(), 21, 128, 50, _stream()))
    assert torch.equal(out, table[ids.view(-1).long()])
    L.check(_lib.lx_embed_rows(table.data_ptr(), ids.data_ptr(), pos.data_ptr(), 7, out.data_ptr(), 21, 128, 50, _stream()))
    assert torch.equal(out, (table[ids.view(-1).long()].float() + pos.repeat(3, 1).float()).to(torch.bfloat16))
    a = torch.randn(9, 64, generator=g, device="cuda").to(torch.bfloat16)
    b = torch.randn(9, 64, generator=g, device="cuda").to(torch.bfloat16)
    want = (a.float() * b.float()).to(torch.bfloat16)
    L.check(_lib.lx_mul_rows(a.data_ptr(), 64, b.data_ptr(), 64, a.data_ptr(), 64, 9, 64, _stream()))  # in place
    assert torch.equal(a, want)


@pytest.mark.parametrize("S", [64, 512])
def test_t5_encoder_matches_the_oracle(S):
    from loongx_b200.text import NativeT5Encoder, T5Config
    from oracle import text_encoders as T

    kw = dict(vocab_size=300, d_model=256, d_kv=64, num_heads=4, d_ff=512, num_layers=3)
    P = _round(T.t5_init(T.T5Cfg(**kw), seed=5))
    enc = NativeT5Encoder(T5Config(**kw), P, "cuda")
    ids = torch.randint(0, 300, (2, S), generator=torch.Generator().manual_seed(S))
    got = enc(ids.cuda())[0]
    with torch.no_grad():
        want = T.t5_encode(P, ids, T.T5Cfg(**kw))
    assert got.shape == (2, S, 256) and got.dtype == torch.bfloat16
    e = _rel(got, want)
    print(f"\n[t5 S={S}] relL2 vs oracle {e:.3g}")
    assert e <= 2e-2


def test_clip_text_matches_the_oracle():
    from loongx_b200.text import ClipTextConfig, NativeClipText
    from oracle import text_encoders as T

    for eos in (2, 290):
        kw = dict(vocab_size=300, hidden_size=128, intermediate_size=256, num_layers=3, num_heads=2, max_positions=77,
                  eos_token_id=eos)
        P = _round(T.clip_init(T.ClipCfg(**kw), seed=6))
        enc = NativeClipText(ClipTextConfig(**kw), P, "cuda")
        ids = torch.randint(3, 280, (3, 77), generator=torch.Generator().manual_seed(2))
        ids[0, 10], ids[1, 76], ids[2, 1] = 290, 290, 290
        o = enc(ids.cuda())
        with torch.no_grad():
            h, pooled = T.clip_encode(P, ids, T.ClipCfg(**kw))
        e1, e2 = _rel(o.last_hidden_state, h), _rel(o.pooler_output, pooled)
        print(f"\n[clip eos={eos}] relL2 hidden {e1:.3g}, pooled {e2:.3g}")
        assert o.pooler_output.shape == (3, 128) and e1 <= 2e-2 and e2 <= 2e-2


def test_generate_from_a_text_prompt():
    """generate(prompt=[...]) with encoders attached = encode_prompt -> (CLIP pooled, T5 hidden states) -> the same
    denoising as passing those embeddings explicitly (generate.py:156-165)."""
    from loongx_b200.config import FluxConfig
    from loongx_b200.text import ClipTextConfig, NativeClipText, NativeT5Encoder, T5Config
    from oracle import sampler as OS
    from oracle import text_encoders as T
    from src.flux.generate import generate
    from src.flux.pipeline_tools import prepare_text_input
    from src.train.model import OminiModel

    cfg = FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2)  # FLUX widths: 4096-wide T5, 768-wide CLIP
    model = OminiModel(cfg, lora_config={"r": 4, "lora_alpha": 4}, device="cuda", model_config={})
    pipe = model.flux_pipe
    with pytest.raises(NotImplementedError, match="attach_text_encoders"):
        generate(model, pipe, prompt="a cat", height=64, width=64, num_inference_steps=1, output_type="latent",
                 use_brain_condition=False)
    tk = dict(vocab_size=300, d_model=4096, d_kv=64, num_heads=4, d_ff=512, num_layers=2)
    ck = dict(vocab_size=300, hidden_size=768, intermediate_size=256, num_layers=2, num_heads=12, max_positions=77)
    PT, PC = _round(T.t5_init(T.T5Cfg(**tk), 7)), _round(T.clip_init(T.ClipCfg(**ck), 8))

    def fake_tokenizer(prompts, padding, max_length, truncation, return_tensors, **kw):
        rows = [[(ord(c) * 7 + i) % 280 + 3 for i, c in enumerate(p[:max_length - 1])] + [299] for p in prompts]
        return {"input_ids": torch.tensor([r + [1] * (max_length - len(r)) for r in rows])}

    pipe.attach_text_encoders(clip=NativeClipText(ClipTextConfig(**ck), PC, "cuda"), t5=NativeT5Encoder(T5Config(**tk), PT, "cuda"),
                              tokenizers=(fake_tokenizer, fake_tokenizer))
    prompts = ["make the sky purple", "add a red hat"]
    pe, po, text_ids = prepare_text_input(pipe, prompts, max_sequence_length=128)
    assert pe.shape == (2, 128, 4096) and po.shape == (2, 768) and text_ids.shape == (128, 3)
    ids_t5 = fake_tokenizer(prompts, "max_length", 128, True, "pt")["input_ids"]
    ids_clip = fake_tokenizer(prompts, "max_length", 77, True, "pt")["input_ids"]
    with torch.no_grad():
        want_pe = T.t5_encode(PT, ids_t5, T.T5Cfg(**tk))
        want_po = T.clip_encode(PC, ids_clip, T.ClipCfg(**ck))[1]
    assert _rel(pe, want_pe) <= 2e-2 and _rel(po, want_po) <= 2e-2
    lat0 = OS.pack_latents(torch.randn(2, 16, 8, 8, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16).cuda())
    kw = dict(height=64, width=64, num_inference_steps=2, output_type="latent", use_brain_condition=False,
              max_sequence_length=128)
    a = generate(model, pipe, prompt=prompts, latents=lat0.clone(), **kw).images
    b = generate(model, pipe, prompt_embeds=pe, pooled_prompt_embeds=po, latents=lat0.clone(), **kw).images
    assert a.shape == (2, 16, 64) and torch.equal(a, b)
