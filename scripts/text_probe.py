"""Time the native text encoders at FLUX.1-dev's sizes (T5 v1.1 XXL, 512 tokens; CLIP-L, 77 tokens) with seeded random
weights created on the device.  Usage: python scripts/text_probe.py [batch]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loongx_b200 import _lib as L  # noqa: E402
from loongx_b200.text import ClipTextConfig, NativeClipText, NativeT5Encoder, T5Config  # noqa: E402


def rnd(g, *shape, std=1.0):
    return (torch.randn(*shape, generator=g, device="cuda", dtype=torch.float32) * std).to(torch.bfloat16)


def t5_params(cfg, g):
    inner = cfg.num_heads * cfg.d_kv
    P = {"encoder.embed_tokens.weight": rnd(g, cfg.vocab_size, cfg.d_model),
         "encoder.final_layer_norm.weight": torch.ones(cfg.d_model, device="cuda"),
         "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight": rnd(g, cfg.num_buckets, cfg.num_heads, std=0.5)}
    for i in range(cfg.num_layers):
        p = f"encoder.block.{i}.layer."
        for n, s in (("q", (cfg.d_model * cfg.d_kv) ** -0.5), ("k", cfg.d_model ** -0.5), ("v", cfg.d_model ** -0.5)):
            P[p + f"0.SelfAttention.{n}.weight"] = rnd(g, inner, cfg.d_model, std=s)
        P[p + "0.SelfAttention.o.weight"] = rnd(g, cfg.d_model, inner, std=inner ** -0.5)
        P[p + "0.layer_norm.weight"] = torch.ones(cfg.d_model, device="cuda")
        P[p + "1.layer_norm.weight"] = torch.ones(cfg.d_model, device="cuda")
        P[p + "1.DenseReluDense.wi_0.weight"] = rnd(g, cfg.d_ff, cfg.d_model, std=cfg.d_model ** -0.5)
        P[p + "1.DenseReluDense.wi_1.weight"] = rnd(g, cfg.d_ff, cfg.d_model, std=cfg.d_model ** -0.5)
        P[p + "1.DenseReluDense.wo.weight"] = rnd(g, cfg.d_model, cfg.d_ff, std=cfg.d_ff ** -0.5)
    return P


def clip_params(cfg, g):
    D, Fd, t = cfg.hidden_size, cfg.intermediate_size, "text_model."
    P = {t + "embeddings.token_embedding.weight": rnd(g, cfg.vocab_size, D, std=0.5),
         t + "embeddings.position_embedding.weight": rnd(g, cfg.max_positions, D, std=0.5),
         t + "final_layer_norm.weight": torch.ones(D, device="cuda"), t + "final_layer_norm.bias": torch.zeros(D, device="cuda")}
    for i in range(cfg.num_layers):
        p = f"{t}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            P[p + f"self_attn.{n}.weight"], P[p + f"self_attn.{n}.bias"] = rnd(g, D, D, std=D ** -0.5), torch.zeros(D, device="cuda")
        for n in ("layer_norm1", "layer_norm2"):
            P[p + n + ".weight"], P[p + n + ".bias"] = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
        P[p + "mlp.fc1.weight"], P[p + "mlp.fc1.bias"] = rnd(g, Fd, D, std=D ** -0.5), torch.zeros(Fd, device="cuda")
        P[p + "mlp.fc2.weight"], P[p + "mlp.fc2.bias"] = rnd(g, D, Fd, std=Fd ** -0.5), torch.zeros(D, device="cuda")
    return P


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def profile(fn):
    L.lib.lx_profile_begin()
    fn()
    t, n, w = (C.c_double * 4)(), (C.c_int64 * 4)(), (C.c_double * 4)()
    L.lib.lx_profile_end(t, n, w)
    return (f"GEMM {t[0]:.2f} ms ({w[0] / max(t[0], 1e-9) / 1e9:.0f} TFLOP/s, {n[0]} launches), attention {t[1]:.2f} ms "
            f"({n[1]} launches), row kernels {t[2]:.2f} ms ({n[2]} launches)")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    g = torch.Generator(device="cuda").manual_seed(0)
    tcfg, ccfg = T5Config(), ClipTextConfig()
    t5 = NativeT5Encoder(tcfg, t5_params(tcfg, g), "cuda")
    clip = NativeClipText(ccfg, clip_params(ccfg, g), "cuda")
    ids5 = torch.randint(0, tcfg.vocab_size, (B, 512), generator=g, device="cuda")
    idsc = torch.randint(0, ccfg.vocab_size, (B, 77), generator=g, device="cuda")
    out = t5(ids5)[0]
    assert torch.isfinite(out.float()).all()
    print(f"[text t5-xxl] B={B} 512 tokens: {timed(lambda: t5(ids5)):.2f} ms/call; {profile(lambda: t5(ids5))}; "
          f"HBM {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    print(f"[text clip-l] B={B} 77 tokens: {timed(lambda: clip(idsc)):.2f} ms/call; {profile(lambda: clip(idsc))}")


if __name__ == "__main__":
    main()
