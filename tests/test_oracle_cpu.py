"""CPU tests that pin the ORACLE: closed-form known answers for the third-party maths it restates (diffusers / s4torch
are absent, SURVEY.md §8c "self-made pins"), the committed golden fixtures, and internal cross-checks."""
import math
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden as MG  # noqa: E402

from oracle import cs3_dgf as OC  # noqa: E402
from oracle import flux_dit as O  # noqa: E402
from oracle import sampler as OS  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def test_sigma_schedule_closed_form_and_golden():
    for n in (4, 28, 50):
        for L in (1024, 4096):
            s = OS.flow_match_sigmas(n, L).double()
            mu = (L - 256) * (1.15 - 0.5) / (4096 - 256) + 0.5
            k = torch.arange(n, dtype=torch.float64)
            lin = 1.0 - k * (1 - 1 / n) / (n - 1)
            ref = math.exp(mu) / (math.exp(mu) + (1 / lin - 1))
            assert torch.allclose(s[:-1], ref, atol=1e-6) and s[-1] == 0 and s[0] == 1.0
            assert np.array_equal(s.float().numpy(), GOLD[f"sigmas_n{n}_L{L}"])
    assert abs(OS.calculate_shift(1024, 256, 4096, 0.5, 1.15) - 0.63) < 1e-9
    assert abs(OS.calculate_shift(4096, 256, 4096, 0.5, 1.15) - 1.15) < 1e-12


def test_rope_tables_closed_form():
    ids = torch.tensor([[0.0, 3.0, 5.0], [0.0, 0.0, 0.0], [0.0, 63.0, 31.0]])
    cos, sin = O.rope_tables(ids)
    assert cos.shape == (3, 128) and cos.dtype == torch.float32
    assert torch.all(cos[:, :16] == 1) and torch.all(sin[:, :16] == 0)  # axis 0 id is always 0
    for r, (a, b) in enumerate([(3.0, 5.0), (0.0, 0.0), (63.0, 31.0)]):
        for j in range(28):
            f = 1.0 / (10000.0 ** (2 * j / 56))
            assert abs(cos[r, 16 + 2 * j].item() - math.cos(a * f)) < 1e-6
            assert cos[r, 16 + 2 * j] == cos[r, 16 + 2 * j + 1]  # repeat_interleave(2)
            assert abs(sin[r, 72 + 2 * j].item() - math.sin(b * f)) < 1e-6
    x = torch.randn(1, 2, 3, 128)
    y = O.apply_rotary_emb(x, (cos, sin))
    xr, xi = x[..., 0::2], x[..., 1::2]
    c2, s2 = cos[:, 0::2], sin[:, 0::2]
    assert torch.allclose(y[..., 0::2], xr * c2 - xi * s2, atol=1e-6)
    assert torch.allclose(y[..., 1::2], xi * c2 + xr * s2, atol=1e-6)
    assert torch.allclose(y.norm(dim=-1), x.norm(dim=-1), atol=1e-5)  # rotations preserve the norm


def test_timestep_embedding_known_values():
    e = O.timestep_sinusoid(torch.tensor([0.0, 500.0, 1000.0]))
    assert e.shape == (3, 256)
    assert torch.all(e[0, :128] == 1) and torch.all(e[0, 128:] == 0)  # cos(0), sin(0): flip_sin_to_cos
    assert abs(e[1, 0].item() - math.cos(500.0)) < 1e-4 and abs(e[1, 128].item() - math.sin(500.0)) < 1e-4
    f127 = math.exp(-math.log(10000.0) * 127 / 128)
    assert abs(e[2, 127].item() - math.cos(1000.0 * f127)) < 1e-5


def test_pack_unpack_ids_bit_exact():
    x = torch.arange(2 * 16 * 8 * 12, dtype=torch.float32).reshape(2, 16, 8, 12)
    p = OS.pack_latents(x)
    assert p.shape == (2, 24, 64)
    assert p[0, 0].tolist()[:4] == [x[0, 0, 0, 0].item(), x[0, 0, 0, 1].item(), x[0, 0, 1, 0].item(), x[0, 0, 1, 1].item()]
    assert torch.equal(OS.unpack_latents(p, 64, 96), x)
    ids = OS.prepare_latent_image_ids(8, 12)
    assert ids.shape == (24, 3) and ids[7].tolist() == [0.0, 1.0, 1.0] and ids[-1].tolist() == [0.0, 3.0, 5.0]
    c = OS.condition_ids(ids, [0, -6], 2.0)
    assert c[7].tolist() == [0.0, 2.5, -9.5] and ids[7].tolist() == [0.0, 1.0, 1.0]


def test_s4_kernel_equals_recurrence():
    """The only available cross-check of SURVEY.md App. B: Cauchy/iFFT kernel == bilinear-discretised recurrence."""
    torch.manual_seed(0)
    n, d, L = 8, 3, 64
    lay = OC.S4Layer(d, n, L)
    lam, p, q = OC.make_nplr(n)
    B, Ct = lay.B.detach().to(torch.complex128), lay.Ct.detach().to(torch.complex128)
    K = OC.s4_kernel(lam, p, q, B, Ct, lay.log_step.detach().double(), L)
    A = torch.diag(lam) - p[:, None] * q.conj()[None, :]
    eye = torch.eye(n, dtype=torch.complex128)
    for c in range(d):
        step = math.exp(lay.log_step[c].item())
        bl = torch.linalg.inv(eye - (step / 2) * A)
        Ab, Bb = bl @ (eye + (step / 2) * A), (bl * step) @ B[c]
        Cc = Ct[c].conj() @ torch.linalg.inv(eye - torch.linalg.matrix_power(Ab, L))
        x = Bb.clone()
        for l in range(L):
            assert abs((Cc @ x).real.item() - K[c, l].item()) < 1e-9 * max(1.0, K.abs().max().item())
            x = Ab @ x
    # float32 module (FFT convolution) vs direct causal convolution with its own kernel
    u = torch.randn(2, L, d)
    y = lay(u)
    Kf = lay.kernel(L)
    direct = torch.stack([sum(Kf[:, j] * u[:, l - j] for j in range(l + 1)) for l in range(L)], 1) + lay.D * u
    assert torch.allclose(y, direct, rtol=1e-4, atol=1e-5)


def test_duan_mask_count_and_shapes():
    torch.manual_seed(1)
    d = OC.DUAN(512).eval()
    x, c = torch.randn(1, 512, 64), torch.randn(1, 512, 64)
    with torch.no_grad():
        y, imp, mask = d(x, c, return_aux=True)
    assert int(mask.sum()) == 358 and int((y.abs().sum(-1) == 0).sum()) == 154  # C - max(1, int(0.7 C)) zeroed
    d1 = OC.DUAN(1).eval()
    with torch.no_grad():
        y1 = d1(torch.randn(2, 1, 768), torch.randn(2, 1, 768))
    assert y1.shape == (2, 1, 768) and (y1.abs().sum(-1) > 0).all()  # C = 1 keeps its only channel


def test_cs3_shapes_and_param_counts():
    m = OC.NeuralConditioner().eval()
    counts = {n: sum(p.numel() for p in getattr(m, n).parameters()) for n in
              ("eeg_projection", "ppg_projection", "fnirs_projection", "motion_projection")}
    # EEG 42.0 M, PPG 6.1 M, fNIRS 6.1 M, Motion 1.1 M (S4 stub excluded there)
    assert 41.9e6 < counts["eeg_projection"] < 42.1e6 and 6.0e6 < counts["ppg_projection"] < 6.2e6
    assert 6.0e6 < counts["fnirs_projection"] < 6.2e6 and 1.0e6 < counts["motion_projection"] < 1.2e6
    with torch.no_grad():
        assert m.ppg_projection(torch.randn(1, 4, 256)).shape == (1, 512, 4096)
        assert m.motion_projection(torch.randn(2, 6, 128)).shape == (2, 768)
    x = torch.randn(1, 4, 100)
    assert torch.equal(OC.spatial_pyramid_pooling(x, 128)[..., :100], x) and OC.spatial_pyramid_pooling(x, 128)[..., 100:].abs().sum() == 0
    assert torch.equal(OC.spatial_pyramid_pooling(x, 64), x[..., :64])


def test_golden_fixtures_reproduce():
    cfg, P, inp = MG.tiny_dit_case()
    with torch.no_grad():
        out64 = MG.tiny_dit_forward(cfg, P, inp, torch.float64)
        out32 = MG.tiny_dit_forward(cfg, P, inp, torch.float32)
    g = torch.from_numpy(GOLD["dit_out64"])
    assert (out64 - g).abs().max() < 1e-9
    assert ((out32.double() - g).norm() / g.norm()) < 1e-5  # fp32 error budget of the tiny DiT
    s4, duan, u, x, c = MG.small_cs3_case()
    with torch.no_grad():
        assert np.allclose(s4(u).numpy(), GOLD["s4_out"], rtol=1e-5, atol=1e-6)
        y, _, mask = duan(x, c, return_aux=True)
    assert np.allclose(y.numpy(), GOLD["duan_out"], rtol=1e-5, atol=1e-6) and np.array_equal(mask.numpy(), GOLD["duan_mask"])


def test_lora_masking_semantics():
    """enable_lora(..., activated=False) == LoRA on the condition rows only (lora_controller.py:5-43)."""
    cfg = O.FluxConfig(**MG.TINY)
    P = O.init_params(cfg, seed=3, lora_b_std=0.1)
    x = torch.randn(2, 5, cfg.inner_dim)
    name = "transformer_blocks.0.attn.to_q"
    base = O.linear(P, name, x, False, cfg)
    lora = O.linear(P, name, x, True, cfg)
    W = P[name + ".weight"] + P[name + ".lora_B.weight"] @ P[name + ".lora_A.weight"] * (cfg.lora_alpha / cfg.lora_rank)
    assert torch.allclose(lora, torch.nn.functional.linear(x, W, P[name + ".bias"]), atol=1e-5)
    assert not torch.allclose(base, lora)
    assert "transformer_blocks.0.attn.add_q_proj.lora_A.weight" not in P  # text stream is never a LoRA target
    assert "transformer_blocks.0.ff.net.0.proj.lora_A.weight" not in P


# ---------------------------------------------------------------------------------------------------------------------
# the DiT restatement pinned against an executable third-party implementation of FLUX (Black Forest Labs' model as
# shipped in the image's torchtitan package): tests/golden/make_dit_bfl_golden.py
# ---------------------------------------------------------------------------------------------------------------------
def _dit_bfl_tools():
    import importlib.util
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_dit_bfl_golden.py")
    spec = importlib.util.spec_from_file_location("make_dit_bfl_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("case", ["plain", "condition"])
def test_oracle_dit_vs_bfl_flux_fixture(case):
    """tests/golden/dit_bfl_v1.npz holds the BFL model's outputs on seeded tiny weights / inputs.  `plain`: the stock
    forward (no condition branch).  `condition`: the reference's three-stream forward (transformer.py:47-252) with the
    condition tokens at c_t = t and LoRA B = 0, where they are arithmetically just more image tokens."""
    import numpy as np

    from oracle import flux_dit as O

    G = _dit_bfl_tools()
    cfg = O.FluxConfig(**G.TINY)
    P = G.params(cfg)
    with_cond = case == "condition"
    got = G.run_oracle(P, cfg, G.inputs(cfg, with_cond=with_cond), with_cond)
    ref = torch.from_numpy(np.load(G.FIXTURE)["out_" + case])
    err = float((got - ref).norm() / ref.norm())
    print(f"[oracle DiT vs BFL fixture, {case}] relL2 {err:.3g}")
    assert err < 1e-5


def test_oracle_dit_vs_bfl_flux_live():
    """The same comparison executed live where torchtitan imports (it is part of this image), on other seeds / shapes and
    with FLUX.1-dev's head count per width ratio (4 heads of 128)."""
    pytest.importorskip("torchtitan.experiments.flux.model.model")
    from oracle import flux_dit as O

    G = _dit_bfl_tools()
    cfg = O.FluxConfig(**dict(G.TINY, num_attention_heads=4, num_layers=1, num_single_layers=2))
    P = G.params(cfg, seed=22)
    for with_cond in (False, True):
        d = G.inputs(cfg, seed=4, B=1, n_txt=16, hw=(4, 10), with_cond=with_cond)
        ref, got = G.run_bfl(P, cfg, d, with_cond), G.run_oracle(P, cfg, d, with_cond)
        err = float((got - ref).norm() / ref.norm())
        print(f"[oracle DiT vs BFL live, condition={with_cond}] relL2 {err:.3g}")
        assert err < 1e-5


def test_oracle_sigma_schedule_vs_bfl_get_schedule():
    """diffusers' FlowMatchEulerDiscreteScheduler.set_timesteps(sigmas=linspace(1, 1/n, n), mu=calculate_shift(L)) as
    restated in oracle/sampler.py against Black Forest Labs' own `get_schedule` (torchtitan's copy of flux/sampling.py):
    the two discretisations coincide (t_i = 1 - i/n, exponential time shift, final 0)."""
    S = pytest.importorskip("torchtitan.experiments.flux.sampling")
    from oracle import sampler as OS

    for n, L in ((4, 1024), (28, 1024), (50, 4096), (28, 256), (8, 2304)):
        a, b = OS.flow_match_sigmas(n, L), torch.tensor(S.get_schedule(n, L))
        assert a.shape == b.shape and float((a - b).abs().max()) < 1e-6, (n, L)
