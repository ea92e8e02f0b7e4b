"""The text-encoder oracle (oracle/text_encoders.py) pinned against the third-party implementation the reference calls
through FluxPipeline.encode_prompt: `transformers`' T5EncoderModel and CLIPTextModel, built here with seeded random
weights (no checkpoints in this image).  Tolerance: fp32 rounding (1e-5 relative)."""
import pytest
import torch

from oracle import text_encoders as T

transformers = pytest.importorskip("transformers")


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def test_t5_oracle_matches_transformers():
    from transformers import T5Config, T5EncoderModel

    cfg = T.T5Cfg(vocab_size=97, d_model=64, d_kv=16, num_heads=4, d_ff=96, num_layers=3)
    hf = T5EncoderModel(T5Config(vocab_size=cfg.vocab_size, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff,
                                 num_layers=cfg.num_layers, num_heads=cfg.num_heads, relative_attention_num_buckets=32,
                                 relative_attention_max_distance=128, feed_forward_proj="gated-gelu", dropout_rate=0.0,
                                 layer_norm_epsilon=1e-6, is_encoder_decoder=False, use_cache=False)).eval()
    P = T.t5_init(cfg, seed=3)
    sd = {k: v for k, v in P.items()}
    sd["shared.weight"] = P["encoder.embed_tokens.weight"]
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("relative_attention_bias" not in m for m in missing), (missing, unexpected)
    ids = torch.randint(0, cfg.vocab_size, (2, 200), generator=torch.Generator().manual_seed(0))  # > max_distance apart
    with torch.no_grad():
        want = hf(input_ids=ids).last_hidden_state
        got = T.t5_encode(P, ids, cfg)
    assert got.shape == (2, 200, 64) and _rel(got, want) < 1e-5, _rel(got, want)


def test_t5_relative_buckets_closed_form():
    b = T.t5_relative_buckets(300)
    assert b.shape == (300, 300) and b.min() == 0 and b.max() == 31
    assert b[5, 5] == 0 and b[5, 4] == 1 and b[5, 6] == 17            # |distance| < 8 is exact; keys after the query: +16
    assert b[0, 7] == 16 + 7 and b[0, 8] == 16 + 8 and b[200, 0] == 15 and b[0, 299] == 31   # log-spaced, capped


def test_clip_oracle_matches_transformers():
    from transformers import CLIPTextConfig, CLIPTextModel

    for eos in (2, 96):  # the legacy argmax pooling and the explicit-EOS pooling
        cfg = T.ClipCfg(vocab_size=97, hidden_size=64, intermediate_size=128, num_layers=3, num_heads=4, max_positions=20,
                        eos_token_id=eos)
        hf = CLIPTextModel(CLIPTextConfig(vocab_size=97, hidden_size=64, intermediate_size=128, num_hidden_layers=3,
                                          num_attention_heads=4, max_position_embeddings=20, hidden_act="quick_gelu",
                                          layer_norm_eps=1e-5, eos_token_id=eos, bos_token_id=0, pad_token_id=1,
                                          attention_dropout=0.0)).eval()
        P = T.clip_init(cfg, seed=4)
        missing, unexpected = hf.load_state_dict(P, strict=False)
        assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)
        ids = torch.randint(3, 90, (3, 20), generator=torch.Generator().manual_seed(1))
        ids[0, 7], ids[1, 19], ids[2, 3] = 96, 96, 96  # EOS (also the largest id) at different positions
        with torch.no_grad():
            o = hf(input_ids=ids)
            h, pooled = T.clip_encode(P, ids, cfg)
        assert _rel(h, o.last_hidden_state) < 1e-5 and _rel(pooled, o.pooler_output) < 1e-5


def test_host_side_of_the_native_encoders():
    """Integer bookkeeping and argument validation of loongx_b200/text.py (no compute without a GPU)."""
    import ctypes as C

    from loongx_b200 import _lib as L
    from loongx_b200 import text as N

    for S in (1, 77, 300):
        assert torch.equal(N.t5_relative_buckets(S, 32, 128), T.t5_relative_buckets(S, 32, 128))
    assert N.T5Config.from_json({"vocab_size": 32128, "d_model": 4096, "d_kv": 64, "num_heads": 64, "d_ff": 10240,
                                 "num_layers": 24, "feed_forward_proj": "gated-gelu"}) == N.T5Config()
    assert N.ClipTextConfig.from_json({"vocab_size": 49408, "hidden_size": 768, "intermediate_size": 3072,
                                       "num_hidden_layers": 12, "num_attention_heads": 12, "hidden_act": "quick_gelu",
                                       "eos_token_id": 2}) == N.ClipTextConfig()
    with pytest.raises(NotImplementedError):
        N.T5Config.from_json({"vocab_size": 1, "d_model": 1, "d_kv": 1, "num_heads": 1, "d_ff": 1, "num_layers": 1,
                              "feed_forward_proj": "relu"})
    lib, one = L.lib, C.c_void_p(16)
    d = N.SmallAttnDesc()
    d.q = d.k = d.v = d.out = 16
    d.ldq = d.ldk = d.ldv = d.ldo = 192
    d.B, d.H, d.S, d.head_dim, d.scale = 1, 1, 77, 32, 1.0
    assert lib.lx_attention_small(C.byref(d), None) != 0 and b"head_dim" in lib.lx_last_error()
    d.head_dim, d.S = 64, 513
    assert lib.lx_attention_small(C.byref(d), None) != 0 and b"S=513" in lib.lx_last_error()
    assert lib.lx_norm_rows(one, 64, one, one, one, 64, 4, 64, 1e-6, 1, None) != 0  # RMS form with a bias
    assert lib.lx_norm_rows(one, 60, one, None, one, 64, 4, 64, 1e-6, 1, None) != 0  # ldx < D
    assert lib.lx_embed_rows(one, one, one, 0, one, 4, 64, 100, None) != 0  # position table without a period
    assert lib.lx_mul_rows(one, 64, one, 64, one, 64, 4, 60, None) != 0  # cols not a multiple of 8

    class _P:  # a pipeline without encoders refuses a text prompt loudly
        text_encoder = text_encoder_2 = tokenizer = tokenizer_2 = None

    from loongx_b200.pipeline import NativeFluxPipeline

    with pytest.raises(NotImplementedError, match="attach_text_encoders"):
        NativeFluxPipeline.encode_prompt(_P(), prompt="a photo")


def test_distance_table_equals_the_full_bias():
    """NativeT5Encoder keeps the relative position bias as [H, 2S-1] indexed by key - query + S - 1 (what
    lx_attention_small's bias_relative mode reads): same values as transformers' full [H, S, S] bias."""
    from loongx_b200.text import t5_relative_buckets

    for S in (1, 2, 77, 300):
        b = t5_relative_buckets(S, 32, 128)
        by_distance = torch.cat([b[1:, 0].flip(0), b[0, :]])
        w = torch.randn(32, 4, generator=torch.Generator().manual_seed(S))
        table = w[by_distance].t()
        idx = torch.arange(S)[None, :] - torch.arange(S)[:, None] + S - 1
        assert torch.equal(table[:, idx], w[T.t5_relative_buckets(S)].permute(2, 0, 1))


def test_text_encoder_checkpoint_directories(tmp_path):
    """<dir>/text_encoder (single safetensors file) and <dir>/text_encoder_2 (sharded, with index) in the layout
    FluxPipeline.from_pretrained reads -> packed native weights (host-side packing only; no kernels run on the CPU)."""
    import json

    from safetensors.torch import save_file

    from loongx_b200 import text as N

    tcfg = T.T5Cfg(vocab_size=50, d_model=128, d_kv=64, num_heads=2, d_ff=256, num_layers=2)
    ccfg = T.ClipCfg(vocab_size=50, hidden_size=128, intermediate_size=256, num_layers=2, num_heads=2, max_positions=77)
    PT, PC = T.t5_init(tcfg, 1), T.clip_init(ccfg, 2)
    PT["shared.weight"] = PT.pop("encoder.embed_tokens.weight")  # real T5 checkpoints may hold only the shared table
    PC["text_model.embeddings.position_ids"] = torch.arange(77)[None]  # legacy buffer in older CLIP checkpoints: ignored
    d1, d2 = tmp_path / "text_encoder", tmp_path / "text_encoder_2"
    d1.mkdir()
    d2.mkdir()
    (d1 / "config.json").write_text(json.dumps({"vocab_size": 50, "hidden_size": 128, "intermediate_size": 256,
                                                "num_hidden_layers": 2, "num_attention_heads": 2, "hidden_act": "quick_gelu",
                                                "max_position_embeddings": 77, "eos_token_id": 2}))
    save_file({k: v.contiguous() for k, v in PC.items()}, str(d1 / "model.safetensors"))
    (d2 / "config.json").write_text(json.dumps({"vocab_size": 50, "d_model": 128, "d_kv": 64, "num_heads": 2, "d_ff": 256,
                                                "num_layers": 2, "feed_forward_proj": "gated-gelu"}))
    keys = sorted(PT)
    shards = {"model-00001-of-00002.safetensors": keys[: len(keys) // 2], "model-00002-of-00002.safetensors": keys[len(keys) // 2:]}
    for fn, ks in shards.items():
        save_file({k: PT[k].contiguous() for k in ks}, str(d2 / fn))
    (d2 / "model.safetensors.index.json").write_text(json.dumps({"weight_map": {k: fn for fn, ks in shards.items() for k in ks}}))
    clip, t5 = N.load_text_encoders(str(tmp_path), device="cpu")
    assert t5.cfg == N.T5Config(50, 128, 64, 2, 256, 2) and clip.cfg.hidden_size == 128
    assert torch.equal(t5.embed, PT["shared.weight"].to(torch.bfloat16))
    a = "encoder.block.1.layer.0.SelfAttention."
    assert torch.equal(t5.layers[1]["qkv"][128:256], PT[a + "k.weight"].to(torch.bfloat16))
    assert t5.layers[0]["wi"].shape == (512, 128) and t5.rel_bias.shape == (32, 2)
    p = "text_model.encoder.layers.0."
    assert torch.equal(clip.layers[0]["qkv_b"][256:], PC[p + "self_attn.v_proj.bias"])
    # quick_gelu folding: fc1 carries the factor 1.702, fc2 its inverse
    assert torch.allclose(clip.layers[0]["fc1_b"], PC[p + "mlp.fc1.bias"] * 1.702)
    assert torch.allclose(clip.layers[0]["fc2"].float(), (PC[p + "mlp.fc2.weight"] / 1.702).to(torch.bfloat16).float())
