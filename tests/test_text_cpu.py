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
