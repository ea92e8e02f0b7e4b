"""CPU checks of the VAE row: the oracle's closed-form pins, the host-side packing / panel arithmetic
of loongx_b200/vae.py against torch convolutions, and the C ABI's argument validation (no compute without a GPU)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from oracle import vae as O


def test_flux_vae_architecture_pins():
    cfg = O.VaeConfig()
    P = O.init_params(cfg)
    # FLUX.1-dev's AutoencoderKL holds 83,819,683 parameters; 244 tensors in its state dict
    assert sum(v.numel() for v in P.values()) == 83_819_683
    assert len(P) == 244
    shapes = O.conv_shapes(cfg)
    assert shapes["decoder.conv_in"] == (512, 16, 3) and shapes["encoder.conv_out"] == (32, 512, 3)
    assert shapes["decoder.up_blocks.2.resnets.0.conv_shortcut"] == (256, 512, 1)
    assert shapes["decoder.up_blocks.3.resnets.0.conv_shortcut"] == (128, 256, 1)
    assert shapes["encoder.down_blocks.1.resnets.0.conv_shortcut"] == (256, 128, 1)
    assert "decoder.up_blocks.3.upsamplers.0.conv" not in shapes and "encoder.down_blocks.3.downsamplers.0.conv" not in shapes
    assert sum(1 for k in shapes if k.startswith("decoder.") and k.endswith(".conv1")) == 2 + 4 * 3


def test_oracle_shapes_and_distribution():
    cfg = O.VaeConfig()
    P = O.init_params(cfg)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(1, 16, 4, 6, generator=g)
    img = O.decode(P, z, cfg)
    assert img.shape == (1, 3, 32, 48) and torch.isfinite(img).all()
    m = O.encode_moments(P, torch.rand(1, 3, 32, 48, generator=g) * 2 - 1, cfg)
    assert m.shape == (1, 32, 4, 6)
    eps = torch.randn(1, 16, 4, 6, generator=g)
    mean, logvar = m.chunk(2, 1)
    assert torch.equal(O.sample_latents(m, None), mean)
    assert torch.allclose(O.sample_latents(m, eps), mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * eps)
    lat = O.encode(P, torch.zeros(1, 3, 32, 48), cfg)
    assert torch.allclose(lat, (O.encode_moments(P, torch.zeros(1, 3, 32, 48), cfg)[:, :16] - 0.1159) * 0.3611)
    # image processor: normalize / denormalize are inverses on [0, 1]; uint8 quantisation rounds to nearest
    x = torch.rand(2, 3, 8, 8, generator=g)
    assert torch.allclose(O.postprocess_pt(O.preprocess(x)), x, atol=1e-6)
    assert O.postprocess_uint8(torch.tensor([[[[-1.0]], [[0.0]], [[1.0]]]])).flatten().tolist() == [0, 128, 255]


def test_host_tables_match_the_oracle():
    from loongx_b200 import vae as V

    ocfg, vcfg = O.VaeConfig(), V.VaeConfig()
    assert V.conv_table(vcfg) == O.conv_shapes(ocfg)
    assert V.norm_table(vcfg) == O.norm_shapes(ocfg)
    po, pv = O.init_params(ocfg, 7), V.synthetic_params(vcfg, 7)
    assert po.keys() == pv.keys() == V.expected_keys(vcfg)
    assert all(torch.equal(po[k], pv[k]) for k in po)
    small = V.VaeConfig(block_out_channels=(32, 64), layers_per_block=1)
    assert V.conv_table(small) == O.conv_shapes(O.VaeConfig(block_out_channels=(32, 64), layers_per_block=1))
    with pytest.raises(NotImplementedError):
        V.VaeConfig.from_json({"use_post_quant_conv": True})
    cj = {"block_out_channels": [128, 256, 512, 512], "latent_channels": 16, "scaling_factor": 0.3611, "shift_factor": 0.1159,
          "use_quant_conv": False, "_class_name": "AutoencoderKL"}
    assert V.VaeConfig.from_json(cj) == vcfg


def _panel_reference(x, taps, up, stride, pad_lo, ldk):
    """The panel lx_vae_im2col is specified to write (include/loongx_b200.h), by plain loops over a small NHWC input."""
    B, H, W, Cc = x.shape
    if taps == 1:
        Ho, Wo = H, W
    elif stride == 1:
        Ho, Wo = H * up, W * up
    else:
        Ho, Wo = (H + 1 - 3) // 2 + 1, (W + 1 - 3) // 2 + 1
    out = torch.zeros(B, Ho, Wo, ldk)
    for oy in range(Ho):
        for ox in range(Wo):
            for tap in range(taps):
                ky, kx = (tap // 3, tap % 3) if taps == 9 else (0, 0)
                iy, ix = oy * stride + ky - pad_lo, ox * stride + kx - pad_lo
                if 0 <= iy < H * up and 0 <= ix < W * up:
                    out[:, oy, ox, tap * Cc:(tap + 1) * Cc] = x[:, iy // up, ix // up]
    return out.reshape(B * Ho * Wo, ldk), Ho, Wo


@pytest.mark.parametrize("case", ["same", "up", "down", "odd_down", "small_cin"])
def test_panel_times_packed_weight_is_the_convolution(case):
    """pack_conv's column order (ky, kx, c) and the panel geometry of the three convolution flavours reproduce
    F.conv2d / Upsample2D / Downsample2D(padding=0)."""
    from loongx_b200.vae import pack_conv

    g = torch.Generator().manual_seed(3)
    cin, cout, H, W = (3, 5, 6, 4) if case == "small_cin" else (8, 16, 6 if case != "odd_down" else 7, 4)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g)
    b = torch.randn(cout, generator=g)
    panel_w, bias, taps, cin_p = pack_conv(w, b)
    assert taps == 9 and cin_p == 8 and panel_w.shape == (16 if cout == 16 else 8, 128) and panel_w.dtype == torch.bfloat16
    wq = panel_w.float()  # compare against the convolution with the same bf16-rounded weights
    w_used = wq[:cout, :9 * cin_p].reshape(cout, 3, 3, cin_p)[..., :cin].permute(0, 3, 1, 2)
    up, stride, pad_lo = (2, 1, 1) if case == "up" else (1, 2, 0) if case in ("down", "odd_down") else (1, 1, 1)
    xr = torch.zeros(2, H, W, cin_p)
    xr[..., :cin] = x.permute(0, 2, 3, 1)
    A, Ho, Wo = _panel_reference(xr, 9, up, stride, pad_lo, panel_w.shape[1])
    got = (A @ wq.t() + bias)[:, :cout].reshape(2, Ho, Wo, cout).permute(0, 3, 1, 2)
    if case == "up":
        want = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w_used, b, padding=1)
    elif stride == 2:
        want = F.conv2d(F.pad(x, (0, 1, 0, 1)), w_used, b, stride=2)
    else:
        want = F.conv2d(x, w_used, b, padding=1)
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=1e-4, rtol=1e-4)


def test_linear_and_shortcut_packing():
    from loongx_b200.vae import pack_conv

    g = torch.Generator().manual_seed(4)
    w, b = torch.randn(24, 64, generator=g), torch.randn(24, generator=g)
    p, bias, taps, cin_p = pack_conv(w, b)
    assert taps == 1 and cin_p == 64 and p.shape == (24, 64)
    assert torch.equal(p, w.to(torch.bfloat16)) and torch.equal(bias, b)
    p2, _, taps2, _ = pack_conv(w[:, :, None, None], b)
    assert taps2 == 1 and torch.equal(p2, p)


def test_vae_entry_points_validate_arguments_without_gpu():
    from loongx_b200 import _lib as L
    from loongx_b200.vae import Im2colDesc

    lib = L.lib
    one = C.c_void_p(16)
    assert lib.lx_vae_group_norm_coeffs(one, 1, 64, 24, 32, one, one, 1e-6, one, one, None) != 0  # C not 8 * 2^k
    assert b"lx_vae_group_norm_coeffs" in lib.lx_last_error()
    assert lib.lx_vae_group_norm_coeffs(one, 1, 64, 128, 48, one, one, 1e-6, one, one, None) != 0  # groups do not divide C
    d = Im2colDesc()
    d.x, d.out = 16, 16
    d.B, d.H, d.W, d.C, d.upsample, d.stride, d.pad_lo, d.taps, d.Ho, d.Wo, d.ldk = 1, 4, 4, 8, 1, 1, 1, 9, 4, 4, 72
    for field, bad in (("C", 12), ("taps", 4), ("upsample", 3), ("stride", 3), ("ldk", 64), ("ldk", 76), ("Ho", 9)):
        keep = getattr(d, field)
        setattr(d, field, bad)
        assert lib.lx_vae_im2col(C.byref(d), None) != 0, field
        assert b"lx_vae_im2col" in lib.lx_last_error()
        setattr(d, field, keep)
    d.taps, d.stride = 1, 2
    assert lib.lx_vae_im2col(C.byref(d), None) != 0
    assert lib.lx_vae_softmax_rows(one, 8, one, 4, 2, 8, 1.0, None) != 0  # ldp < n
    assert lib.lx_vae_nchw_to_rows(one, one, 1, 3, 16, 4, 1.0, 0.0, None) != 0  # c_pad not a multiple of 8
    assert lib.lx_vae_rows_to_nchw(one, 2, one, 1, 3, 16, 0, None) != 0  # ld < C
    assert lib.lx_vae_sample_latents(one, 16, None, one, 1, 16, 4, 0.0, 1.0, None) != 0  # ld < 2 L


def test_image_processor_preprocess_on_host():
    from loongx_b200.vae import ImageProcessor

    ip = ImageProcessor(16)
    x = torch.rand(1, 3, 16, 16)
    assert torch.allclose(ip.preprocess(x), 2 * x - 1)
    assert torch.equal(ip.preprocess(2 * x - 1), 2 * x - 1)  # already normalised: passed through
    PIL = pytest.importorskip("PIL.Image")
    im = PIL.new("RGB", (32, 16), (255, 0, 128))
    t = ip.preprocess(im)
    assert t.shape == (1, 3, 16, 32) and torch.allclose(t[0, :, 0, 0], torch.tensor([1.0, -1.0, 2 * 128 / 255 - 1]), atol=1e-6)
    assert ip.preprocess(PIL.new("RGB", (30, 16))).shape == (1, 3, 16, 16)  # resized down to the 16-pixel grid
    with pytest.raises(ValueError):
        ip.preprocess(PIL.new("RGB", (30, 8)))
    assert ip.postprocess(x, "latent") is x


def test_pipeline_without_vae_raises_loudly():
    from src.flux.pipeline_tools import encode_images

    class _P:
        vae = None
        device, dtype = "cpu", torch.float32

    with pytest.raises(NotImplementedError, match="attach_vae"):
        encode_images(_P(), torch.zeros(1, 3, 32, 32))


def test_vae_checkpoint_directory_round_trip(tmp_path):
    """<dir>/vae in the diffusers layout -> VaeWeights: config, key validation and panels (host-side packing runs on any
    device; no kernels are involved)."""
    from loongx_b200 import vae as V

    cfg = V.VaeConfig(block_out_channels=(32, 64), layers_per_block=1, scaling_factor=0.5, shift_factor=0.25)
    P = V.synthetic_params(cfg, 3)
    V.write_diffusers_vae(str(tmp_path), cfg, P)
    w = V.VaeWeights.from_pretrained(str(tmp_path), device="cpu")
    assert w.cfg == cfg
    direct = V.VaeWeights(cfg, P, "cpu")
    assert w.conv.keys() == direct.conv.keys() and w.norm.keys() == direct.norm.keys()
    for k in w.conv:
        assert torch.equal(w.conv[k].w, direct.conv[k].w) and torch.equal(w.conv[k].bias, direct.conv[k].bias)
    qkv = w.conv["decoder.mid_block.attentions.0.to_qkv"]
    assert qkv.w.shape == (3 * 64, 64) and qkv.taps == 1
    assert torch.equal(qkv.w[64:128], P["decoder.mid_block.attentions.0.to_k.weight"].to(torch.bfloat16))
    bad = dict(P)
    bad.pop("decoder.conv_out.bias")
    with pytest.raises(KeyError, match="missing"):
        V.VaeWeights(cfg, bad, "cpu")
    bad = dict(P)
    bad["decoder.conv_in.weight"] = bad["decoder.conv_in.weight"][:, :8]
    with pytest.raises(ValueError, match="decoder.conv_in"):
        V.VaeWeights(cfg, bad, "cpu")


def test_encode_images_ids_fallback_on_host():
    """encode_images' id construction incl. the reference's retry with the halved grid (pipeline_tools.py:15-29), with a
    torch-only pipeline stand-in (latents in, so no VAE and no kernels are involved)."""
    from oracle import sampler as OS
    from src.flux.pipeline_tools import encode_images

    class _Pipe:
        device, dtype, vae = "cpu", torch.float32, None
        _pack_latents = staticmethod(lambda lat, *a: OS.pack_latents(lat))
        _prepare_latent_image_ids = staticmethod(lambda B, h, w, dev, dt: OS.prepare_latent_image_ids(h, w).to(dt))

    class _Legacy(_Pipe):  # a diffusers flavour that counts the id grid per latent pixel
        _prepare_latent_image_ids = staticmethod(lambda B, h, w, dev, dt: torch.zeros(h * w, 3))

    z = torch.randn(2, 16, 8, 12, generator=torch.Generator().manual_seed(0))
    tokens, ids = encode_images(_Pipe(), z)
    assert torch.equal(tokens, OS.pack_latents(z)) and torch.equal(ids, OS.prepare_latent_image_ids(8, 12))
    tokens, ids = encode_images(_Legacy(), z)
    assert ids.shape == (24, 3)


# ---------------------------------------------------------------------------------------------------------------------
# the oracle pinned against an executable third-party implementation of the FLUX.1 autoencoder
# ---------------------------------------------------------------------------------------------------------------------
def _bfl_tools():
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_vae_bfl_golden.py")
    spec = importlib.util.spec_from_file_location("make_vae_bfl_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _oracle_three(P, cfg, images, latents, seed):
    from oracle import vae as V

    moments = V.encode_moments(P, images, cfg)
    torch.manual_seed(seed)  # the BFL module draws randn_like(mean) inside encode(): the same draw here
    eps = torch.randn(moments.shape[0], cfg.latent_channels, *moments.shape[2:])  # (contiguous, like the module's chunk view)
    return moments, V.encode(P, images, cfg, eps=eps), V.decode(P, latents, cfg)


def test_oracle_vae_vs_bfl_autoencoder_fixture():
    """tests/golden/vae_bfl_v1.npz: outputs of Black Forest Labs' AutoEncoder (torchtitan's copy) on seeded weights /
    inputs at reduced width (generated by tests/golden/make_vae_bfl_golden.py).  The oracle must reproduce them: encoder
    moments, the sampled + shifted + scaled latents, the decoded image."""
    import numpy as np

    from oracle import vae as V

    G = _bfl_tools()
    d = np.load(G.FIXTURE)
    cfg = V.VaeConfig(**G.SMALL)
    P = V.init_params(cfg, seed=77)
    images, latents = torch.from_numpy(d["images"]), torch.from_numpy(d["latents"])
    moments, z, img = _oracle_three(P, cfg, images, latents, seed=5)
    for name, got in (("moments", moments), ("z", z), ("img", img)):
        ref = torch.from_numpy(d[name])
        err = float((got - ref).norm() / ref.norm())
        print(f"[oracle vs BFL fixture] {name}: relL2 {err:.3g}")
        assert err < 1e-5, name


@pytest.mark.parametrize("width", ["small", "flux"])
def test_oracle_vae_vs_bfl_autoencoder_live(width):
    """The same comparison executed live where the torchtitan package is importable (it is part of this image), at the
    fixture's width and at FLUX.1-dev's real widths (128, 256, 512, 512)."""
    pytest.importorskip("torchtitan.experiments.flux.model.autoencoder")
    from oracle import vae as V

    G = _bfl_tools()
    cfg = V.VaeConfig(**G.SMALL) if width == "small" else V.VaeConfig()
    P = V.init_params(cfg, seed=78)
    images, latents = G.inputs(seed=12, hw=(40, 56) if width == "small" else (32, 48))
    ref = G.run_bfl(P, cfg, images, latents, seed=6)
    got = _oracle_three(P, cfg, images, latents, seed=6)
    for name, a, b in zip(("moments", "z", "img"), got, ref):
        err = float((a - b).norm() / b.norm())
        print(f"[oracle vs BFL live, {width}] {name}: relL2 {err:.3g}")
        assert err < 1e-5, name
