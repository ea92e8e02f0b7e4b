"""GPU test of the checkpoint path: a diffusers-format transformer directory + a peft LoRA file +
a LoongX full state dict, loaded through the reference's own entry points (`OminiModel(flux_pipe_id=dir)`,
`model.load_lora`, `model.load_state_dict`), must give the same DiT forward as the oracle with those parameters."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def _forward(model, inp, B):
    from loongx_b200.dit import DitPlan

    plan = DitPlan(model.transformer.weights, B, 128, 128, 128, T=1)
    plan.set_ids(inp["txt_ids"], inp["img_ids"], inp["cond_ids"])
    plan.prepare(inp["pe"], inp["pooled"], inp["cond"], [inp["t"]] * B, [3.5] * B, c_t=0.0)
    out = plan.step(0, inp["lat"])
    torch.cuda.synchronize()
    return out


def test_checkpoint_formats_end_to_end(tmp_path):
    from safetensors.torch import save_file

    from oracle import flux_dit as O
    from oracle import sampler as OS
    from loongx_b200 import checkpoint as CK
    from loongx_b200.config import FluxConfig
    from src.train.model import OminiModel

    dev = "cuda"
    kw = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
    ocfg, cfg = O.FluxConfig(**kw), FluxConfig(**kw)
    P = {k: v.to(torch.bfloat16) for k, v in O.init_params(ocfg, seed=5, w_std=0.05, bias_std=0.05, lora_b_std=0.05).items()}
    base = {k: v for k, v in P.items() if ".lora_" not in k}
    lora = {k: v.float() for k, v in P.items() if ".lora_" in k}
    CK.write_diffusers_transformer(str(tmp_path / "flux"), cfg, base, shards=2)
    save_file({"transformer." + k: v for k, v in lora.items()}, str(tmp_path / "pytorch_lora_weights.safetensors"))

    g = torch.Generator().manual_seed(9)
    B = 1
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).bfloat16().to(dev)  # noqa: E731
    img_ids = OS.prepare_latent_image_ids(16, 32).to(dev)
    inp = dict(lat=r(B, 128, 64), cond=r(B, 128, 64), pe=r(B, 128, 256, scale=0.5), pooled=r(B, 64), img_ids=img_ids,
               cond_ids=OS.condition_ids(img_ids, [0, -16]), txt_ids=torch.zeros(128, 3, device=dev), t=0.6)

    def oracle(params):
        P32 = {k: v.float().to(dev) for k, v in params.items()}
        with torch.no_grad():
            return O.tranformer_forward(P32, ocfg, inp["cond"].float(), inp["cond_ids"], None, {}, 0, hidden_states=inp["lat"].float(),
                                        encoder_hidden_states=inp["pe"].float(), pooled_projections=inp["pooled"].float(),
                                        timestep=torch.full((B,), inp["t"], device=dev), img_ids=inp["img_ids"],
                                        txt_ids=inp["txt_ids"], guidance=torch.full((B,), 3.5, device=dev))

    # 1. diffusers directory: fresh LoRA (B = 0) -> forward == base model with LoRA disabled on every branch
    m = OminiModel(str(tmp_path / "flux"), lora_config={"r": 4, "lora_alpha": 4}, device=dev)
    assert m.transformer.cfg.num_layers == 1 and m.transformer.cfg.inner_dim == 256
    e0 = _rel(_forward(m, inp, B), oracle(base))
    # 2. peft LoRA file -> native re-merge
    assert m.load_lora(str(tmp_path)) is m  # model.py:464-477 returns self
    e1 = _rel(_forward(m, inp, B), oracle(P))
    d01 = _rel(oracle(P), oracle(base))
    # 3. LoongX full state dict with the peft-injected spellings, different LoRA factors
    P2 = dict(P)
    for k in lora:
        P2[k] = (lora[k] * 1.5).bfloat16()
    sd = {}
    targets = {k.rsplit(".lora_", 1)[0] for k in lora}
    for k, v in P2.items():
        stem, kind = k.rsplit(".", 1)
        if ".lora_" in k:
            mod, ab = k[:-len(".weight")].rsplit(".lora_", 1)
            sd[f"transformer.{mod}.lora_{ab}.default.weight"] = v
        elif stem in targets:
            sd[f"transformer.{stem}.base_layer.{kind}"] = v
        else:
            sd["transformer." + k] = v
    sd.update({k: v.clone() for k, v in m.state_dict().items() if not k.startswith("transformer.")})  # CS3 / DGF, strict
    res = m.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    e2 = _rel(_forward(m, inp, B), oracle(P2))
    # 4. export: state_dict() in the LoongX layout and save_lora() round-trip bit-exactly through a second model
    sd_out = m.state_dict()
    assert any(k.endswith("attn.to_q.base_layer.weight") for k in sd_out) and "transformer.proj_out.weight" in sd_out
    m2 = OminiModel(str(tmp_path / "flux"), lora_config={"r": 4, "lora_alpha": 4}, device=dev)
    m2.load_state_dict(sd_out)
    assert torch.equal(_forward(m2, inp, B), _forward(m, inp, B))
    m.save_lora(str(tmp_path / "exported"))
    m3 = OminiModel(str(tmp_path / "flux"), lora_config={"r": 4, "lora_alpha": 4}, device=dev)
    m3.load_lora(str(tmp_path / "exported"))
    assert torch.equal(_forward(m3, inp, B), _forward(m, inp, B))
    print(f"\n[checkpoint] relL2 vs oracle: directory {e0:.4g}, +LoRA file {e1:.4g}, LoongX state dict {e2:.4g}; LoRA effect {d01:.4g}")
    assert max(e0, e1, e2) < 2e-2 and d01 > 5 * max(e0, e1)
    # 5. a FLUX directory that also holds vae/ (FluxPipeline.from_pretrained loads it, model.py:398-400): the pipeline gets
    #    the native VAE and image processor, with the file's scaling / shift factors
    from loongx_b200.vae import VaeConfig, synthetic_params, write_diffusers_vae
    from oracle import vae as OV

    assert m.flux_pipe.vae is None
    vcfg = VaeConfig(scaling_factor=0.5, shift_factor=0.25)
    VP = synthetic_params(vcfg, 5)
    write_diffusers_vae(str(tmp_path / "flux"), vcfg, VP)
    m4 = OminiModel(str(tmp_path / "flux"), lora_config={"r": 4, "lora_alpha": 4}, device=dev)
    pipe = m4.flux_pipe
    assert pipe.vae is not None and pipe.image_processor is not None and pipe.vae_scale_factor == 16
    assert (pipe.vae.config.scaling_factor, pipe.vae.config.shift_factor) == (0.5, 0.25)
    z = torch.randn(1, 16, 8, 8, generator=torch.Generator().manual_seed(9))
    got = pipe.vae.decode(z.to(dev), return_dict=False)[0]
    want = OV.decode_raw(VP, z, OV.VaeConfig())
    assert _rel(got.cpu(), want) < 3e-2
