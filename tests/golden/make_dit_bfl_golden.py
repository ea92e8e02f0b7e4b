"""Pins oracle/flux_dit.py -- the restatement of diffusers' FluxTransformer2DModel arithmetic that the reference's
block.py / transformer.py call into -- against an independent, executable implementation of the same network: Black Forest
Labs' FLUX model as shipped in the `torchtitan` package of this image (torchtitan/experiments/flux/model/{model,layers,
math}.py: EmbedND RoPE, timestep_embedding, MLPEmbedder, Modulation, QK-RMSNorm, DoubleStreamBlock, SingleStreamBlock,
LastLayer).  torchtitan's copy has no guidance embedder (it is the MLPEmbedder once more), so the comparison runs with
`guidance_embeds=False`.

Two cases, both through the oracle's `tranformer_forward` (the reference's transformer.py:47-252 restated):
  plain      no condition branch: the stock FLUX forward
  condition  the reference's three-stream forward with the condition tokens at c_t = t and LoRA B = 0: then the condition
             tokens are arithmetically just more image tokens (same norm1 / attention / feed-forward weights, same
             conditioning vector), so BFL's model fed img = [image tokens ; condition tokens] must give the same image rows

`to_bfl_state_dict` is the published key correspondence between the diffusers layout (oracle) and the BFL layout
(diffusers' convert_flux_transformer_checkpoint_to_diffusers run backwards, incl. the shift / scale swap of norm_out).

  python tests/golden/make_dit_bfl_golden.py        # writes tests/golden/dit_bfl_v1.npz (tiny widths, seeded)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import flux_dit as O  # noqa: E402

FIXTURE = os.path.join(ROOT, "tests", "golden", "dit_bfl_v1.npz")
TINY = dict(num_layers=2, num_single_layers=3, num_attention_heads=2, joint_attention_dim=96, pooled_projection_dim=48,
            guidance_embeds=False)


def to_bfl_state_dict(P, cfg):
    D = cfg.inner_dim
    out = {}

    def lin(dst, *srcs):  # one BFL Linear = the row-concatenation of one or more diffusers Linears
        out[dst + ".weight"] = torch.cat([P[s + ".weight"] for s in srcs], 0).clone()
        out[dst + ".bias"] = torch.cat([P[s + ".bias"] for s in srcs], 0).clone()

    lin("img_in", "x_embedder")
    lin("txt_in", "context_embedder")
    lin("time_in.in_layer", "time_text_embed.timestep_embedder.linear_1")
    lin("time_in.out_layer", "time_text_embed.timestep_embedder.linear_2")
    lin("vector_in.in_layer", "time_text_embed.text_embedder.linear_1")
    lin("vector_in.out_layer", "time_text_embed.text_embedder.linear_2")
    for i in range(cfg.num_layers):
        s, d = f"transformer_blocks.{i}.", f"double_blocks.{i}."
        lin(d + "img_mod.lin", s + "norm1.linear")
        lin(d + "txt_mod.lin", s + "norm1_context.linear")
        lin(d + "img_attn.qkv", s + "attn.to_q", s + "attn.to_k", s + "attn.to_v")
        lin(d + "txt_attn.qkv", s + "attn.add_q_proj", s + "attn.add_k_proj", s + "attn.add_v_proj")
        out[d + "img_attn.norm.query_norm.weight"] = P[s + "attn.norm_q.weight"].clone()
        out[d + "img_attn.norm.key_norm.weight"] = P[s + "attn.norm_k.weight"].clone()
        out[d + "txt_attn.norm.query_norm.weight"] = P[s + "attn.norm_added_q.weight"].clone()
        out[d + "txt_attn.norm.key_norm.weight"] = P[s + "attn.norm_added_k.weight"].clone()
        lin(d + "img_attn.proj", s + "attn.to_out.0")
        lin(d + "txt_attn.proj", s + "attn.to_add_out")
        lin(d + "img_mlp.0", s + "ff.net.0.proj")
        lin(d + "img_mlp.2", s + "ff.net.2")
        lin(d + "txt_mlp.0", s + "ff_context.net.0.proj")
        lin(d + "txt_mlp.2", s + "ff_context.net.2")
    for i in range(cfg.num_single_layers):
        s, d = f"single_transformer_blocks.{i}.", f"single_blocks.{i}."
        lin(d + "modulation.lin", s + "norm.linear")
        lin(d + "linear1", s + "attn.to_q", s + "attn.to_k", s + "attn.to_v", s + "proj_mlp")
        lin(d + "linear2", s + "proj_out")
        out[d + "norm.query_norm.weight"] = P[s + "attn.norm_q.weight"].clone()
        out[d + "norm.key_norm.weight"] = P[s + "attn.norm_k.weight"].clone()
    lin("final_layer.linear", "proj_out")
    w, b = P["norm_out.linear.weight"], P["norm_out.linear.bias"]  # diffusers: (scale | shift); BFL: (shift | scale)
    out["final_layer.adaLN_modulation.1.weight"] = torch.cat([w[D:], w[:D]], 0).clone()
    out["final_layer.adaLN_modulation.1.bias"] = torch.cat([b[D:], b[:D]], 0).clone()
    return out


def bfl_model(P, cfg):
    from torchtitan.experiments.flux.model.args import FluxModelArgs
    from torchtitan.experiments.flux.model.model import FluxModel

    m = FluxModel(FluxModelArgs(in_channels=cfg.in_channels, out_channels=cfg.in_channels, vec_in_dim=cfg.pooled_projection_dim,
                                context_in_dim=cfg.joint_attention_dim, hidden_size=cfg.inner_dim, mlp_ratio=float(cfg.mlp_ratio),
                                num_heads=cfg.num_attention_heads, depth=cfg.num_layers,
                                depth_single_blocks=cfg.num_single_layers, axes_dim=tuple(cfg.axes_dims_rope), theta=10_000,
                                qkv_bias=True)).float().eval()
    m.load_state_dict(to_bfl_state_dict(P, cfg), strict=True)
    for mod in m.modules():  # BFL's own RMSNorm (and diffusers') use eps 1e-6; torchtitan's nn.RMSNorm defaults to finfo.eps
        if isinstance(mod, torch.nn.RMSNorm):
            mod.eps = 1e-6
    return m


def inputs(cfg, seed=3, B=2, n_txt=24, hw=(6, 8), with_cond=True):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, scale=1.0: torch.randn(*s, generator=g) * scale  # noqa: E731
    h, w = hw
    n_img = h * w
    ids = torch.zeros(h, w, 3)
    ids[..., 1] += torch.arange(h)[:, None]
    ids[..., 2] += torch.arange(w)[None, :]
    img_ids = ids.reshape(n_img, 3)
    cond_ids = img_ids.clone()
    cond_ids[:, 2] -= w  # the reference's position_delta = (0, -w)
    d = dict(hidden_states=r(B, n_img, cfg.in_channels), encoder_hidden_states=r(B, n_txt, cfg.joint_attention_dim, scale=0.5),
             pooled_projections=r(B, cfg.pooled_projection_dim), timestep=torch.full((B,), 0.37), img_ids=img_ids,
             txt_ids=torch.zeros(n_txt, 3))
    if with_cond:
        d.update(condition_latents=r(B, n_img, cfg.in_channels), condition_ids=cond_ids)
    return d


def run_oracle(P, cfg, d, with_cond):
    kw = {k: d[k] for k in ("hidden_states", "encoder_hidden_states", "pooled_projections", "timestep", "img_ids", "txt_ids")}
    with torch.no_grad():
        if with_cond:  # c_t = t: the condition stream sees the image stream's conditioning vector
            return O.tranformer_forward(P, cfg, d["condition_latents"], d["condition_ids"], None, {}, float(d["timestep"][0]), **kw)
        return O.tranformer_forward(P, cfg, None, None, None, {}, 0, **kw)


def run_bfl(P, cfg, d, with_cond):
    m = bfl_model(P, cfg)
    B = d["hidden_states"].shape[0]
    img, img_ids = d["hidden_states"], d["img_ids"]
    if with_cond:
        img = torch.cat([img, d["condition_latents"]], 1)
        img_ids = torch.cat([img_ids, d["condition_ids"]], 0)
    with torch.no_grad():
        out = m(img=img, img_ids=img_ids[None].expand(B, -1, -1), txt=d["encoder_hidden_states"],
                txt_ids=d["txt_ids"][None].expand(B, -1, -1), timesteps=d["timestep"], y=d["pooled_projections"])
    return out[:, : d["hidden_states"].shape[1]]


def params(cfg, seed=21):
    # biases switched on, LoRA B = 0 (peft's initial state): the adapters contribute nothing, like in the stock model
    return O.init_params(cfg, seed=seed, w_std=0.05, bias_std=0.05, lora_b_std=0.0)


if __name__ == "__main__":
    cfg = O.FluxConfig(**TINY)
    P = params(cfg)
    save = {}
    for name, with_cond in (("plain", False), ("condition", True)):
        d = inputs(cfg, with_cond=with_cond)
        ref = run_bfl(P, cfg, d, with_cond)
        got = run_oracle(P, cfg, d, with_cond)
        print(name, "BFL vs oracle relL2", float((got - ref).norm() / ref.norm()), tuple(ref.shape))
        save["out_" + name] = ref.numpy()
    np.savez_compressed(FIXTURE, **save)
    print(FIXTURE)
