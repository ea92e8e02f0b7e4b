"""GPU parity of the native VAE against oracle/vae.py (the restatement of diffusers 0.31.0's
AutoencoderKL, pinned at 1e-6 to Black Forest Labs' executable FLUX AutoEncoder: tests/test_vae_cpu.py, see the oracle header).

Tolerances: the native path stores activations in bf16 and accumulates in fp32 (GroupNorm statistics in fp64), the
oracle is fp32 throughout.  Kernel-level checks compare against the same arithmetic on the same bf16 inputs (<= 1 bf16
ulp); the 30-layer decoder / encoder are held to relL2 <= 3e-2 and the decoded image to a mean error under 2/255."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def vae():
    from loongx_b200.vae import NativeVae, VaeConfig, VaeWeights, synthetic_params

    cfg = VaeConfig()
    P = synthetic_params(cfg, 1234)
    return NativeVae(VaeWeights(cfg, P, "cuda")), P


def _ocfg():
    from oracle import vae as O

    return O, O.VaeConfig()


def test_group_norm_coefficients(vae):
    v, _ = vae
    from loongx_b200.vae import _Act

    g = torch.Generator(device="cuda").manual_seed(1)
    for (B, H, W, Cc) in ((2, 8, 8, 128), (1, 33, 17, 512), (3, 40, 40, 256)):
        x = (torch.randn(B, H, W, Cc, generator=g, device="cuda") * 2 + 0.5).to(torch.bfloat16)
        name = {128: "decoder.conv_norm_out", 512: "decoder.mid_block.resnets.0.norm1", 256: "decoder.up_blocks.2.resnets.1.norm1"}[Cc]
        gamma, beta = v.w.norm[name]
        coeff = v._coeffs(_Act(x.reshape(-1, Cc), B, H, W), name)
        got = x.float() * coeff[:, None, None, :, 0] + coeff[:, None, None, :, 1]
        want = F.group_norm(x.float().permute(0, 3, 1, 2), 32, gamma, beta, 1e-6).permute(0, 2, 3, 1)
        assert (got - want).abs().max().item() < 2e-4


@pytest.mark.parametrize("gather", [False, True])
@pytest.mark.parametrize("case", ["same", "up", "down", "odd_down", "norm_silu", "one_tap_norm"])
def test_im2col_panels(vae, case, gather):
    """Both forms of the panel kernel (scatter: thread per input pixel, used when ldk == 9 C; gather: thread per panel
    vector) against torch's unfold."""
    from loongx_b200.vae import _lib

    _lib.lx_debug_vae_im2col_gather(int(gather))
    try:
        _check_panel(vae, case)
    finally:
        _lib.lx_debug_vae_im2col_gather(0)


def _check_panel(vae, case):
    v, _ = vae
    from loongx_b200.vae import _Act

    g = torch.Generator(device="cuda").manual_seed(2)
    B, H, W, Cc = 2, (7 if case == "odd_down" else 6), 10, 128
    x = torch.randn(B, H, W, Cc, generator=g, device="cuda").to(torch.bfloat16)
    a = _Act(x.reshape(-1, Cc), B, H, W)
    up, stride, pad_lo = (2, 1, 1) if case == "up" else (1, 2, 0) if case in ("down", "odd_down") else (1, 1, 1)
    taps = 1 if case == "one_tap_norm" else 9
    coeff = v._coeffs(a, "decoder.conv_norm_out") if case in ("norm_silu", "one_tap_norm") else None
    ldk = taps * Cc + (64 if case == "same" else 0)
    panel, Ho, Wo = v._panel(a, ldk, taps, coeff, silu=(case == "norm_silu"), up=up, stride=stride, pad_lo=pad_lo)
    xs = x.float()
    if coeff is not None:
        xs = xs * coeff[:, None, None, :, 0] + coeff[:, None, None, :, 1]
        if case == "norm_silu":
            xs = F.silu(xs)
    xs = xs.permute(0, 3, 1, 2)
    if taps == 1:
        want = xs.permute(0, 2, 3, 1).reshape(B * H * W, Cc)
    else:
        if up == 2:
            xs = F.interpolate(xs, scale_factor=2.0, mode="nearest")
        xs = F.pad(xs, (0, 1, 0, 1)) if stride == 2 else F.pad(xs, (1, 1, 1, 1))
        cols = F.unfold(xs, 3, stride=stride)  # [B, C*9, L] with channel-major (c, ky, kx) ordering
        L = cols.shape[-1]
        want = cols.reshape(B, Cc, 9, L).permute(0, 3, 2, 1).reshape(B * L, 9 * Cc)  # -> (ky, kx, c)
        assert L == Ho * Wo
    assert panel.shape == (want.shape[0], ldk)
    tol = 0.0 if coeff is None else 3e-2  # bf16 rounding of values up to ~6, SiLU through tanh.approx
    assert (panel[:, :taps * Cc].float() - want).abs().max().item() <= tol
    assert (panel[:, taps * Cc:] == 0).all()


def test_softmax_rows_and_layout_kernels(vae):
    from loongx_b200 import _lib as L
    from loongx_b200.vae import _lib, _stream

    g = torch.Generator(device="cuda").manual_seed(3)
    n, ld = 1000, 1008
    s = torch.randn(37, ld, generator=g, device="cuda") * 8
    p = torch.full((37, ld), 7.0, dtype=torch.bfloat16, device="cuda")
    L.check(_lib.lx_vae_softmax_rows(s.data_ptr(), ld, p.data_ptr(), ld, 37, n, 0.25, _stream()))
    want = torch.softmax(s[:, :n] * 0.25, -1)
    assert (p[:, :n].float() - want).abs().max().item() < 4e-3 and (p[:, n:] == 0).all()
    assert (p[:, :n].float().sum(-1) - 1).abs().max().item() < 1e-2
    # NCHW fp32 -> bf16 rows with channel padding and affine, and back
    x = torch.randn(2, 3, 5, 6, generator=g, device="cuda")
    rows = torch.empty(2 * 30, 8, dtype=torch.bfloat16, device="cuda")
    L.check(_lib.lx_vae_nchw_to_rows(x.data_ptr(), rows.data_ptr(), 2, 3, 30, 8, 2.0, -1.0, _stream()))
    want = (2 * x - 1).permute(0, 2, 3, 1).reshape(60, 3).to(torch.bfloat16)
    assert torch.equal(rows[:, :3], want) and (rows[:, 3:] == 0).all()
    r32 = torch.randn(60, 8, generator=g, device="cuda") * 2
    back = torch.empty(2, 3, 5, 6, device="cuda")
    L.check(_lib.lx_vae_rows_to_nchw(r32.data_ptr(), 8, back.data_ptr(), 2, 3, 30, 1, _stream()))
    assert torch.equal(back, (r32[:, :3].reshape(2, 5, 6, 3).permute(0, 3, 1, 2) * 0.5 + 0.5).clamp(0, 1))
    # latent sampling
    mom = torch.randn(60, 32, generator=g, device="cuda")
    mom[:, 16:] *= 20
    eps = torch.randn(2, 16, 5, 6, generator=g, device="cuda")
    out = torch.empty(2, 16, 5, 6, device="cuda")
    L.check(_lib.lx_vae_sample_latents(mom.data_ptr(), 32, eps.data_ptr(), out.data_ptr(), 2, 16, 30, 0.1159, 0.3611, _stream()))
    m = mom.reshape(2, 5, 6, 32).permute(0, 3, 1, 2)
    want = (m[:, :16] + torch.exp(0.5 * m[:, 16:].clamp(-30, 20)) * eps - 0.1159) * 0.3611
    assert torch.allclose(out, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 8, 8), (1, 6, 10), (1, 16, 16)])
def test_decode_matches_the_oracle(vae, shape):
    v, P = vae
    O, ocfg = _ocfg()
    B, h, w = shape
    g = torch.Generator().manual_seed(10 + h)
    z = torch.randn(B, 16, h, w, generator=g)
    want = O.decode_raw(P, z, ocfg)
    got = v.decode(z.cuda(), return_dict=False)[0]
    assert got.shape == want.shape and got.dtype == torch.float32
    print(f"\n[vae decode {shape}] relL2 vs oracle {_rel(got, want):.3g}")
    assert _rel(got, want) <= 3e-2, _rel(got, want)
    img_err = (O.postprocess_pt(got.cpu()) - O.postprocess_pt(want)).abs().mean().item()
    assert img_err < 2 / 255, img_err


@pytest.mark.parametrize("shape", [(2, 64, 64), (1, 48, 80)])
def test_encode_matches_the_oracle(vae, shape):
    v, P = vae
    O, ocfg = _ocfg()
    B, H, W = shape
    g = torch.Generator().manual_seed(20 + H)
    img = torch.rand(B, 3, H, W, generator=g) * 2 - 1
    want_m = O.encode_moments(P, img, ocfg)
    rows, nb, h, w = v.encode_moments(img.cuda())
    got_m = rows.reshape(B, h, w, 32).permute(0, 3, 1, 2)
    assert (nb, h, w) == (B, H // 8, W // 8)
    print(f"\n[vae encode {shape}] relL2 vs oracle {_rel(got_m, want_m):.3g}")
    assert _rel(got_m, want_m) <= 3e-2, _rel(got_m, want_m)
    dist = v.encode(img.cuda()).latent_dist
    assert _rel(dist.mode(), want_m[:, :16]) <= 3e-2
    gen = torch.Generator(device="cuda").manual_seed(5)
    eps = torch.randn(B, 16, h, w, generator=gen, device="cuda")
    gen.manual_seed(5)
    z = dist.sample(gen)
    own_m = dist._m.reshape(B, h, w, 32).permute(0, 3, 1, 2)  # the moments this distribution holds
    want_z = O.sample_latents(own_m.cpu(), eps.cpu())  # same moments, same noise: checks the sampling arithmetic
    assert torch.allclose(z.cpu(), want_z, rtol=1e-4, atol=1e-4)
    assert torch.equal(own_m, got_m)  # two passes over the same input are bit-identical (fixed reduction orders)


def test_batch_split_and_full_size_decode(vae):
    """512 x 512 decode (BASELINE's image size): finite, samples independent of their batch neighbours and of the
    panel-memory split."""
    v, _ = vae
    g = torch.Generator(device="cuda").manual_seed(7)
    z = torch.randn(2, 16, 64, 64, generator=g, device="cuda")
    both = v.decode(z, return_dict=False)[0]
    assert both.shape == (2, 3, 512, 512) and torch.isfinite(both).all()
    keep = v.PANEL_BYTES
    try:
        v.PANEL_BYTES = 1 << 30  # forces one sample per pass
        one = v.decode(z, return_dict=False)[0]
    finally:
        v.PANEL_BYTES = keep
    again = v.decode(z, return_dict=False)[0]
    print(f"\n[vae 512x512] split vs joint {_rel(one, both):.3g}, run to run {_rel(again, both):.3g}")
    # every reduction has a fixed order (no atomics): bit-reproducible run to run, and a sample does not depend on its
    # batch neighbours
    assert torch.equal(again, both)
    assert _rel(one, both) < 1e-3, _rel(one, both)
    assert v.launches > 0


def test_generate_decodes_through_the_pipeline():
    """generate(..., output_type='pt') = unpack -> z / scale + shift -> native decode -> denormalize (generate.py:375-380);
    encode_images() with a VAE attached = normalize -> native encode -> sample -> (z - shift) * scale -> pack."""
    from loongx_b200.config import FluxConfig
    from loongx_b200.vae import VaeConfig, synthetic_params
    from oracle import sampler as OS
    from src.flux.condition import Condition
    from src.flux.generate import generate
    from src.flux.pipeline_tools import encode_images
    from src.train.model import OminiModel

    O, ocfg = _ocfg()
    H_PX, W_PX = 256, 128
    cfg = FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2)
    model = OminiModel(cfg, lora_config={"r": 4, "lora_alpha": 4}, device="cuda", model_config={})
    pipe = model.flux_pipe
    with pytest.raises(NotImplementedError, match="attach_vae"):
        encode_images(pipe, torch.zeros(1, 3, 32, 32))
    pipe.attach_vae(None, seed=1234)
    assert pipe.vae_scale_factor == 16
    P = synthetic_params(VaeConfig(), 1234)
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g, device="cuda")  # noqa: E731
    pe, po = (0.1 * r(1, 512, 4096)).bfloat16(), r(1, 768).bfloat16()
    packed0 = OS.pack_latents(r(1, 16, 32, 16).bfloat16())
    cond = r(1, 16, 32, 16).bfloat16()
    kw = dict(prompt_embeds=pe, pooled_prompt_embeds=po, height=H_PX, width=W_PX, num_inference_steps=2, default_lora=True,
              use_brain_condition=False)

    def run(output_type):
        c = Condition("subject", condition=cond, position_delta=[0, -8])
        return generate(model, pipe, conditions=[c], latents=packed0.clone(), output_type=output_type, **kw).images

    lat = run("latent")
    img = run("pt")
    assert img.shape == (1, 3, H_PX, W_PX) and img.min() >= 0 and img.max() <= 1
    z = pipe._unpack_latents(lat, H_PX, W_PX, pipe.vae_scale_factor).float().cpu()
    want = O.postprocess_pt(O.decode(P, z, ocfg))
    err = (img.cpu() - want).abs().mean().item()
    assert err < 2 / 255, err
    pil = run("pil")
    assert len(pil) == 1 and pil[0].size == (W_PX, H_PX)
    # encode side: a [0, 1] picture -> tokens of its sampled, shifted and scaled latents
    pic = torch.rand(1, 3, 64, 96, generator=torch.Generator().manual_seed(1))
    torch.manual_seed(3)
    tokens, ids = encode_images(pipe, pic)
    assert tokens.shape == (1, 4 * 6, 64) and ids.shape == (4 * 6, 3)
    torch.manual_seed(3)
    eps = torch.randn(1, 16, 8, 12, device="cuda")
    want_lat = O.encode(P, O.preprocess(pic).to(torch.bfloat16).float(), ocfg, eps.cpu())
    want_tokens = pipe._pack_latents(want_lat.cuda().to(pipe.dtype))
    assert _rel(tokens, want_tokens) <= 3e-2, _rel(tokens, want_tokens)


def test_condition_from_a_raw_picture():
    """inference.py:86-98: Condition("subject", raw_img=PIL) -> host preprocessing -> image_processor.preprocess -> native VAE
    encode -> tokens + shifted ids, and generate() runs on it."""
    import numpy as np
    from PIL import Image

    from loongx_b200.config import FluxConfig
    from loongx_b200.vae import VaeConfig, synthetic_params
    from oracle import sampler as OS
    from src.flux.condition import Condition
    from src.flux.generate import generate
    from src.train.model import OminiModel

    O, ocfg = _ocfg()
    cfg = FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2)
    model = OminiModel(cfg, lora_config={"r": 4, "lora_alpha": 4}, device="cuda", model_config={})
    pipe = model.flux_pipe
    pipe.attach_vae(None, seed=1234)
    P = synthetic_params(VaeConfig(), 1234)
    rng = np.random.default_rng(1)
    pic = Image.fromarray(rng.integers(0, 255, (128, 256, 3), dtype=np.uint8))  # 256 wide, 128 high
    cond = Condition("coloring", raw_img=pic, position_delta=[0, -16])
    torch.manual_seed(11)
    tokens, ids, type_id = cond.encode(pipe)
    assert tokens.shape == (1, 8 * 16, 64) and ids.shape == (128, 3) and int(type_id[0]) == 6
    assert torch.equal(ids[:, 2].cpu().float(), (OS.prepare_latent_image_ids(16, 32)[:, 2] - 16).float())
    torch.manual_seed(11)
    eps = torch.randn(1, 16, 16, 32, device="cuda")
    gray = np.asarray(pic.convert("L").convert("RGB"), dtype=np.float32) / 255.0
    x = (2 * torch.from_numpy(gray).permute(2, 0, 1)[None] - 1).to(torch.bfloat16).float()
    want = pipe._pack_latents(O.encode(P, x, ocfg, eps.cpu()).cuda().to(pipe.dtype))
    assert _rel(tokens, want) <= 3e-2, _rel(tokens, want)
    g = torch.Generator(device="cuda").manual_seed(0)
    pe, po = (0.1 * torch.randn(1, 512, 4096, generator=g, device="cuda")).bfloat16(), torch.randn(1, 768, generator=g, device="cuda").bfloat16()
    out = generate(model, pipe, conditions=[cond], prompt_embeds=pe, pooled_prompt_embeds=po, height=128, width=256,
                   num_inference_steps=2, output_type="pil", default_lora=True, use_brain_condition=False,
                   generator=torch.Generator(device="cuda").manual_seed(1)).images
    assert len(out) == 1 and out[0].size == (256, 128)


def test_decode_and_encode_vs_bfl_autoencoder_directly(vae):
    """The native VAE against an executable THIRD-PARTY implementation, not through this repo's oracle: Black Forest Labs'
    FLUX AutoEncoder (torchtitan's copy, fp32 on the CPU) with the same weights mapped by
    tests/golden/make_vae_bfl_golden.py, at FLUX.1-dev's full widths.  Decoder on z / scale + shift, encoder moments."""
    pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location(
        "make_vae_bfl_golden", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_vae_bfl_golden.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    v, P = vae
    O, ocfg = _ocfg()
    ae = G.bfl_autoencoder({k: t.float().cpu() for k, t in P.items()}, ocfg)
    g = torch.Generator().manual_seed(31)
    lat = torch.randn(1, 16, 24, 40, generator=g) * 0.8  # "pipeline" latents: the decoder sees lat / scale + shift
    img = torch.rand(1, 3, 64, 96, generator=g) * 2 - 1
    with torch.no_grad():
        want_img = ae.decode(lat)
        want_m = ae.encoder(img)
    got_img = v.decode((lat / ocfg.scaling_factor + ocfg.shift_factor).cuda(), return_dict=False)[0]
    rows, nb, h, w = v.encode_moments(img.cuda())
    got_m = rows.reshape(1, h, w, 32).permute(0, 3, 1, 2)
    e_d, e_e = _rel(got_img, want_img), _rel(got_m, want_m)
    print(f"\n[native VAE vs BFL AutoEncoder] decode relL2 {e_d:.3g}, encoder moments relL2 {e_e:.3g}")
    assert e_d <= 3e-2 and e_e <= 3e-2
    assert (O.postprocess_pt(got_img.cpu()) - O.postprocess_pt(want_img)).abs().mean().item() < 2 / 255
