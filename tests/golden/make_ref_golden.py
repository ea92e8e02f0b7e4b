"""Generates tests/golden/ref_v1.npz by EXECUTING THE REFERENCE'S OWN SOURCE (/root/reference/src/flux/*.py,
src/train/model.py) on the CPU through oracle/ref_harness.py (third-party packages replaced by stand-ins, see its
header).  Only runs in the build container (the GPU box has no /root/reference); the fixture it writes is committed.

    python tests/golden/make_ref_golden.py

Every case is a seeded function in `CASES`; tests/test_reference_pins_cpu.py replays the same seeded inputs through the
oracle restatement and compares with the stored reference outputs, and (when the tree is present) re-runs the reference
live.  Large outputs are stored as a strided sample + float64 sum / abs-sum so the fixture stays small.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import cs3_dgf as OC  # noqa: E402
from oracle import flux_dit as O  # noqa: E402
from oracle import ref_harness as R  # noqa: E402
from oracle import sampler as OS  # noqa: E402

TINY = dict(num_layers=2, num_single_layers=2, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
TINY_BRAIN = dict(num_layers=1, num_single_layers=1, num_attention_heads=2)  # joint 4096 / pooled 768 like FLUX
DIT_VARIANTS = {
    "default": dict(model_config={}),
    "latent_lora": dict(model_config={"latent_lora": True}),
    "no_union": dict(model_config={"union_cond_attn": False}),
    "independent": dict(model_config={"independent_condition": True}),
    "c_factor": dict(model_config={}, c_factor=1.7),
    "no_cond": dict(model_config={}, use_cond=False),
    "c_t": dict(model_config={}, c_t=0.25),
}


def digest(t: torch.Tensor, max_elems=4096) -> dict:
    """Strided sample + sums of a tensor (what the fixture stores)."""
    f = t.detach().to(torch.float64).flatten()
    step = max(1, f.numel() // max_elems)
    return {"sample": f[::step].to(torch.float32).numpy(), "sum": np.float64(f.sum()), "abssum": np.float64(f.abs().sum()),
            "shape": np.array(t.shape)}


# ------------------------------------------------------------------------------------------------------------------
# seeded inputs shared by generator and tests
# ------------------------------------------------------------------------------------------------------------------
def dit_inputs(cfg, B=2, nt=24, h=8, w=16, seed=1):
    g = torch.Generator().manual_seed(seed)
    ni = (h // 2) * (w // 2)
    img_ids = OS.prepare_latent_image_ids(h, w)
    return dict(lat=torch.randn(B, ni, 64, generator=g), cond=torch.randn(B, ni, 64, generator=g),
                pe=torch.randn(B, nt, cfg.joint_attention_dim, generator=g) * 0.5,
                pooled=torch.randn(B, cfg.pooled_projection_dim, generator=g), img_ids=img_ids,
                cond_ids=OS.condition_ids(img_ids, [0, -(w // 2)]), txt_ids=torch.zeros(nt, 3),
                t=torch.tensor([0.65, 0.3][:B]), guidance=torch.full((B,), 3.5))


def dit_params(cfg):
    return O.init_params(cfg, seed=1234, dtype=torch.float32, w_std=0.05, bias_std=0.05, lora_b_std=0.05)


def signals(seed=45, B=1):
    """SURVEY.md §8d synthetic signals: EEG truncate path, fNIRS truncate, PPG exact, Motion zero-pad."""
    g = torch.Generator().manual_seed(seed)
    return dict(eeg=torch.randn(B, 4, 5000, generator=g), fnirs=torch.randn(B, 6, 600, generator=g),
                ppg=torch.randn(B, 4, 256, generator=g), motion=torch.randn(B, 6, 100, generator=g))


def make_conditioner(seed=1234) -> OC.NeuralConditioner:
    torch.manual_seed(seed)
    return OC.NeuralConditioner().eval()


# ------------------------------------------------------------------------------------------------------------------
# reference-side runners (need /root/reference)
# ------------------------------------------------------------------------------------------------------------------
def ref_dit_forward(cfg, P, inp, model_config, c_factor=None, use_cond=True, c_t=0, **extra):
    T = R.ref_module("flux.transformer")
    model = R.build_transformer(P, cfg)
    if c_factor is not None:  # generate.py:90-94
        for name, m in model.named_modules():
            if name.endswith(".attn"):
                m.c_factor = torch.ones(1, 1) * c_factor
    with torch.no_grad():
        return T.tranformer_forward(
            model, inp["cond"] if use_cond else None, inp["cond_ids"].clone() if use_cond else None, None, model_config, c_t,
            hidden_states=inp["lat"], encoder_hidden_states=inp["pe"], pooled_projections=inp["pooled"], timestep=inp["t"],
            img_ids=inp["img_ids"], txt_ids=inp["txt_ids"], guidance=inp["guidance"], return_dict=False, **extra)[0]


def oracle_dit_forward(cfg, P, inp, model_config, c_factor=None, use_cond=True, c_t=0, **extra):
    with torch.no_grad():
        return O.tranformer_forward(
            P, cfg, inp["cond"] if use_cond else None, inp["cond_ids"] if use_cond else None, None, model_config, c_t,
            hidden_states=inp["lat"], encoder_hidden_states=inp["pe"], pooled_projections=inp["pooled"], timestep=inp["t"],
            img_ids=inp["img_ids"], txt_ids=inp["txt_ids"], guidance=inp["guidance"], c_factor=c_factor, **extra)


class _D1(nn.Module):
    """Deviation D1: generate.py:215-232 hands the encoders `x.flatten(1)`, which crashes in
    EEGEncoder.forward (model.py:76 permutes a 2-D tensor).  This adaptor restores [B, C, L] and changes nothing else."""

    def __init__(self, enc, channels):
        super().__init__()
        self.enc, self.channels = enc, channels

    def forward(self, flat):
        return self.enc(flat.view(flat.shape[0], self.channels, -1))


def ref_omini_model(nc: OC.NeuralConditioner, d1: bool = True):
    """An OminiModel (model.py:376) assembled WITHOUT its constructor (which downloads FLUX): the reference's own
    encoder / DUAN classes and its unbound methods, holding the parameters of the oracle conditioner `nc`."""
    M = R.ref_module("train.model")
    m = M.OminiModel.__new__(M.OminiModel)
    nn.Module.__init__(m)
    m.eeg_fixed_length, m.fnirs_fixed_length, m.ppg_fixed_length, m.motion_fixed_length = 4096, 512, 256, 128
    kw = dict(device="cpu", dtype=torch.float32)
    encs = dict(eeg_projection=(M.EEGEncoder, 4), ppg_projection=(M.PPGEncoder, 4), fnirs_projection=(M.FNIRSEncoder, 6),
                motion_projection=(M.MotionEncoder, 6))
    for name, (cls, ch) in encs.items():
        enc = cls(**kw).eval()
        R.copy_module_params(enc, getattr(nc, name))
        setattr(m, name, _D1(enc, ch) if d1 else enc)
    for name, ch in dict(duan_norm1=512, duan_norm2=1, duan_norm_prompt=512, duan_norm_pooled=1).items():
        d = M.DUAN(channels=ch, **kw).eval()
        R.copy_module_params(d, getattr(nc, name))
        setattr(m, name, d)
    for name in ("fusion1", "fusion2", "fusion3", "fusion4"):
        src = getattr(nc, name)
        lin = nn.Sequential(nn.Linear(src[0].in_features, src[0].out_features))
        R.copy_module_params(lin, src)
        setattr(m, name, lin)
    return m.eval(), M


def ref_generate(cfg, P, inp, nc=None, sig=None, n_steps=4, h=8, w=16, fuse_flag=True, condition_scale=1.0,
                 use_brain_condition=False, model_config=None):
    """The reference's generate() (generate.py:66-394) end to end on the CPU: stand-in pipeline, tiny DiT, packed
    latents in, output_type='latent'.  Condition latents go through Condition.encode -> encode_images with the VAE
    factored out (harness _PrecomputedVae), so the id arithmetic of condition.py:126-137 runs for real."""
    G = R.ref_module("flux.generate")
    Cn = R.ref_module("flux.condition")
    model, _ = ref_omini_model(nc) if nc is not None else (None, None)
    pipe = R.FluxPipeline(R.build_transformer(P, cfg))
    pipe.vae = R._PrecomputedVae(shift=0.0, scale=1.0)
    pipe.image_processor = types.SimpleNamespace(preprocess=lambda z: z)
    B = inp["lat"].shape[0]
    cond_z = OS.unpack_latents(inp["cond"], h * 8, w * 8)  # [B,16,h,w] whose packing is inp["cond"]
    cond = Cn.Condition("subject", condition=cond_z, position_delta=[0, -(w // 2)])
    kw = {}
    if use_brain_condition:
        # generate.py:171 unsqueezes un-batched signals; B=1 here
        kw = dict(additional_condition1=sig["eeg"][0], additional_condition2=sig["fnirs"][0],
                  additional_condition3=sig["ppg"][0], additional_condition4=sig["motion"][0])
    out = G.generate(model, pipe, conditions=[cond], model_config=model_config or {"union_cond_attn": True},
                     condition_scale=condition_scale, default_lora=True, use_brain_condition=use_brain_condition,
                     fuse_flag=fuse_flag, prompt_embeds=inp["pe"], pooled_prompt_embeds=inp["pooled"], height=h * 8,
                     width=w * 8, num_inference_steps=n_steps, latents=inp["lat"], output_type="latent", **kw)
    assert B == out.images.shape[0]
    return out.images


def oracle_generate(cfg, P, inp, nc=None, sig=None, n_steps=4, fuse_flag=True, condition_scale=1.0,
                    use_brain_condition=False, model_config=None):
    pe, pooled = inp["pe"], inp["pooled"]
    if use_brain_condition:
        with torch.no_grad():
            pe, pooled = nc.conditioning(pe, pooled, sig["eeg"], sig["fnirs"], sig["ppg"], sig["motion"], fuse_flag=fuse_flag,
                                         mode="generate")
    return OS.denoise(P, cfg, inp["lat"], pe, pooled, inp["txt_ids"], inp["img_ids"], inp["cond"], inp["cond_ids"],
                      num_inference_steps=n_steps, guidance_scale=3.5, model_config=model_config or {},
                      c_factor=None if condition_scale == 1.0 else condition_scale)


def cs3_cases(nc: OC.NeuralConditioner, sig, ref=None):
    """name -> output tensor for every CS3 / DGF unit; `ref` = (reference OminiModel, module) or None for the oracle."""
    out = {}
    with torch.no_grad():
        if ref is None:
            spp, m = OC.spatial_pyramid_pooling, nc
            enc = lambda name, x: getattr(nc, name)(x)  # noqa: E731
        else:
            m, M = ref
            spp = lambda x, n, adaptive=False: M.OminiModel.spatial_pyramid_pooling(m, x, n, adaptive)  # noqa: E731
            enc = lambda name, x: getattr(m, name).enc(x)  # noqa: E731
        fixed = dict(eeg=4096, fnirs=512, ppg=256, motion=128)
        x = {k: spp(v, fixed[k]) for k, v in sig.items()}
        for k in x:
            out["spp_" + k] = x[k]
        out["spp_adaptive"] = spp(sig["fnirs"], 512, True)
        e = out["enc_eeg"] = enc("eeg_projection", x["eeg"])
        p = out["enc_ppg"] = enc("ppg_projection", x["ppg"])
        f = out["enc_fnirs"] = enc("fnirs_projection", x["fnirs"])
        mo = out["enc_motion"] = enc("motion_projection", x["motion"])
        out["fuse_eeg"] = m.fuse_eeg(e, p)
        out["fuse_fnirs"] = m.fuse_fnirs(f, mo)
        g = torch.Generator().manual_seed(7)
        a, b = torch.randn(2, 512, 96, generator=g), torch.randn(2, 512, 96, generator=g) * 2 + 0.3
        out["duan512"] = m.duan_norm_prompt(a, b)
        out["duan512_keep_half"] = m.duan_norm_prompt(a, b, keep_ratio=0.5)
        out["duan1"] = m.duan_norm_pooled(a[:, :1], b[:, :1])
        fpp_in = torch.randn(2, 3, 300, generator=g)
        if ref is None:
            out["fpp"] = OC.FeaturePyramidPooling([16, 50, 128])(fpp_in)
        else:
            out["fpp"] = ref[1].FeaturePyramidPooling(output_sizes=[16, 50, 128])(fpp_in)
    return out


def condition_id_cases(ref: bool):
    """Condition.encode id arithmetic (condition.py:106-138) incl. position_scale; bit-exact."""
    out = {}
    h, w = 8, 16
    g = torch.Generator().manual_seed(3)
    z = torch.randn(1, 16, h, w, generator=g)
    for name, (delta, scale) in {"delta": ([0, -8], 1.0), "none": (None, 1.0), "scale2": ([2, -3], 2.0)}.items():
        if ref:
            Cn = R.ref_module("flux.condition")
            pipe = R.FluxPipeline(None)
            pipe.vae = R._PrecomputedVae()
            pipe.image_processor = types.SimpleNamespace(preprocess=lambda t: t)
            c = Cn.Condition("canny", condition=z, position_delta=delta, position_scale=scale)
            tokens, ids, type_id = c.encode(pipe)
            out[f"cond_{name}_type"] = type_id
        else:
            tokens = OS.pack_latents((z - 0.1159) * 0.3611)
            ids = OS.condition_ids(OS.prepare_latent_image_ids(h, w), delta, scale)
        out[f"cond_{name}_tokens"], out[f"cond_{name}_ids"] = tokens, ids
    return out


# ------------------------------------------------------------------------------------------------------------------
# training step (model.py:569-729): loss + LoRA gradients
# ------------------------------------------------------------------------------------------------------------------
STEP_CASES = {"text": dict(brain=False, fuse=True, seed=11), "brain_fuse1": dict(brain=True, fuse=True, seed=12),
              "brain_fuse0": dict(brain=True, fuse=False, seed=13)}


def step_batch(cfg, brain: bool, seed: int):
    g = torch.Generator().manual_seed(seed)
    B, h, w = (1, 8, 16) if brain else (2, 8, 16)
    nt = 512 if brain else 24
    batch = dict(image=torch.randn(B, 16, h, w, generator=g), condition=torch.randn(B, 16, h, w, generator=g),
                 prompt_embeds=torch.randn(B, nt, cfg.joint_attention_dim, generator=g) * 0.5,
                 pooled_prompt_embeds=torch.randn(B, cfg.pooled_projection_dim, generator=g),
                 position_delta=[[0, -(w // 2)]], condition_type=["subject"] * B)
    if brain:
        batch.update(signals(seed=seed + 100, B=B))
    return batch


def ref_step(cfg, P, batch, nc, brain: bool, fuse: bool, seed: int):
    """The reference's OminiModel.step + loss.backward() on the CPU (stand-in pipeline / transformer, real step code)."""
    model, M = ref_omini_model(nc, d1=False)
    tr = R.build_transformer(P, cfg)
    for n, p in tr.named_parameters():
        p.requires_grad_(".lora_A." in n or ".lora_B." in n)
    pipe = R.FluxPipeline(tr)
    pipe.vae = R._PrecomputedVae(shift=0.0, scale=1.0)
    pipe.image_processor = types.SimpleNamespace(preprocess=lambda z: z)
    object.__setattr__(model, "flux_pipe", pipe)
    object.__setattr__(model, "transformer", tr)
    model.model_config, model.use_brain_condition, model.fuse_flag = {}, brain, fuse
    model._dtype = torch.float32
    type(model).device = property(lambda self: torch.device("cpu"))
    b = dict(image=batch["image"], condition=batch["condition"], condition_type=batch["condition_type"],
             description=(batch["prompt_embeds"], batch["pooled_prompt_embeds"]), position_delta=batch["position_delta"])
    for k in ("eeg", "fnirs", "ppg", "motion"):
        if k in batch:
            b[k] = batch[k]
    torch.manual_seed(seed)  # t = sigmoid(randn(B)); x_1 = randn_like(x_0)  (model.py:590-591)
    loss = model.step(b)
    loss.backward()
    grads = {}
    for n, p in tr.named_parameters():
        if p.requires_grad:
            key = n.replace(".default.weight", ".weight")
            grads[key] = p.grad if p.grad is not None else torch.zeros_like(p)
    return loss.detach(), grads


def oracle_step(cfg, P, batch, nc, brain: bool, fuse: bool, seed: int):
    from oracle import train_step as TS

    torch.manual_seed(seed)
    loss, grads, _ = TS.flow_step_grads(P, cfg, batch, model_config={}, conditioner=nc, use_brain_condition=brain,
                                        fuse_flag=fuse)
    return loss, grads


def step_cases(ref: bool, nc) -> dict:
    out = {}
    for name, c in STEP_CASES.items():
        cfg = O.FluxConfig(**(TINY_BRAIN if c["brain"] else TINY))
        P = dit_params(cfg)
        batch = step_batch(cfg, c["brain"], c["seed"])
        loss, grads = (ref_step if ref else oracle_step)(cfg, P, batch, nc, c["brain"], c["fuse"], c["seed"])
        out[f"step_{name}_loss"] = loss.reshape(1)
        out[f"step_{name}_grads"] = torch.cat([grads[k].flatten() for k in sorted(grads)])
    return out


GPU_DIT_VARIANTS = {"default": {}, "independent": {"independent_condition": True}, "no_union": {"union_cond_attn": False}}
GPU_DIT_TS = (0.9, 0.35)


def gpu_dit_case(seed=0, B=2):
    """The case tests/test_dit_gpu.py replays through the CUDA path: stream lengths are multiples of 128 tokens (the
    native tiling), weights and inputs are bf16-rounded (what the CUDA path stores) and evaluated here in fp32."""
    cfg = O.FluxConfig(**TINY)
    P = {k: v.to(torch.bfloat16).float() for k, v in dit_params(cfg).items()}
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).bfloat16().float()  # noqa: E731
    img_ids = OS.prepare_latent_image_ids(16, 32)
    inp = dict(lat=r(B, 128, 64), cond=r(B, 128, 64), pe=r(B, 128, 256, scale=0.5), pooled=r(B, 64), img_ids=img_ids,
               cond_ids=OS.condition_ids(img_ids, [0, -16]), txt_ids=torch.zeros(128, 3), guidance=torch.full((B,), 3.5))
    return cfg, P, inp


def gpu_dit_cases(ref: bool) -> dict:
    cfg, P, inp = gpu_dit_case()
    out = {}
    for name, mc in GPU_DIT_VARIANTS.items():
        for s, t in enumerate(GPU_DIT_TS):
            i = dict(inp, t=torch.full((inp["lat"].shape[0],), t))
            out[f"gpudit_{name}_s{s}"] = (ref_dit_forward if ref else oracle_dit_forward)(cfg, P, i, mc)
    return out


def all_cases(ref: bool) -> dict:
    """Every pinned quantity, from the reference (ref=True) or from the oracle restatement (ref=False)."""
    res = {}
    cfg = O.FluxConfig(**TINY)
    P, inp = dit_params(cfg), dit_inputs(cfg)
    for name, kw in DIT_VARIANTS.items():
        res["dit_" + name] = (ref_dit_forward if ref else oracle_dit_forward)(cfg, P, inp, **kw)
    gen = ref_generate if ref else oracle_generate
    res["gen_text4"] = gen(cfg, P, inp, n_steps=4)
    res["gen_text7_cscale"] = gen(cfg, P, inp, n_steps=7, condition_scale=1.3)
    nc, sig = make_conditioner(), signals()
    res.update({"cs3_" + k: v for k, v in cs3_cases(nc, sig, ref_omini_model(nc) if ref else None).items()})
    cfgb = O.FluxConfig(**TINY_BRAIN)
    Pb, inpb = dit_params(cfgb), dit_inputs(cfgb, B=1, nt=512)
    for fuse in (True, False):
        res[f"gen_brain_fuse{int(fuse)}"] = gen(cfgb, Pb, inpb, nc=nc, sig=sig, n_steps=2, fuse_flag=fuse,
                                                use_brain_condition=True)
    res.update(condition_id_cases(ref))
    res.update(gpu_dit_cases(ref))
    res.update(step_cases(ref, nc))
    return res


CONTROLNET_CASES = {"2+2": (2, 2), "1+1": (1, 1), "2+0": (2, 0), "0+1": (0, 1)}


def controlnet_cases(ref: bool) -> dict:
    """tranformer_forward with controlnet residuals (transformer.py:172-181, 230-239): residual lists as long as / shorter
    than the block lists (interval = ceil(n_blocks / n_samples)), either list absent.  Kept in its own fixture
    (ref_controlnet_v1.npz, written by `python make_ref_golden.py controlnet`) so that ref_v1.npz stays byte-identical."""
    cfg = O.FluxConfig(**TINY)
    P, inp = dit_params(cfg), dit_inputs(cfg)
    B, ni = inp["lat"].shape[:2]
    out = {}
    for name, (n_dbl, n_sgl) in CONTROLNET_CASES.items():
        g = torch.Generator().manual_seed(9)
        mk = lambda n: [torch.randn(B, ni, cfg.inner_dim, generator=g) * 0.3 for _ in range(n)] if n else None  # noqa: E731
        extra = dict(controlnet_block_samples=mk(n_dbl), controlnet_single_block_samples=mk(n_sgl))
        out["controlnet_" + name] = (ref_dit_forward if ref else oracle_dit_forward)(cfg, P, inp, {}, **extra)
    return out


def main_controlnet():
    assert R.available(), "needs /root/reference"
    ref, orc = controlnet_cases(True), controlnet_cases(False)
    store = {}
    for k, v in ref.items():
        a, b = v.double(), orc[k].double()
        print(f"{k:28s} shape {tuple(v.shape)!s:20s} relL2(oracle, reference) = {float((a - b).norm() / (a.norm() + 1e-30)):.3e}")
        for kk, vv in digest(v).items():
            store[f"{k}/{kk}"] = vv
    path = os.path.join(HERE, "ref_controlnet_v1.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "controlnet":
        return main_controlnet()
    assert R.available(), "needs /root/reference"
    ref, orc = all_cases(True), all_cases(False)
    store = {}
    for k, v in ref.items():
        if k in orc:
            a, b = v.double(), orc[k].double()
            err = float((a - b).norm() / (a.norm() + 1e-30))
            print(f"{k:28s} shape {tuple(v.shape)!s:20s} relL2(oracle, reference) = {err:.3e}")
        for kk, vv in digest(v).items():
            store[f"{k}/{kk}"] = vv
        if k.startswith("gpudit_"):
            store[f"{k}/full"] = v.numpy()  # whole output: the CUDA parity test compares against it directly
    np.savez_compressed(os.path.join(HERE, "ref_v1.npz"), **store)
    print("wrote", os.path.join(HERE, "ref_v1.npz"), os.path.getsize(os.path.join(HERE, "ref_v1.npz")), "bytes")


if __name__ == "__main__":
    main()
