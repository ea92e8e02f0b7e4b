"""GPU parity of the native CS3 encoders / DGF fusion (fp32 kernels through the C ABI) against the oracle modules
(oracle/cs3_dgf.py) loaded with the SAME state dict.

Stated tolerance: float32 path, relL2 <= 1e-4 (FFT-convolution vs direct-convolution and reduction-order
differences only); DUAN top-k channel mask identical (importance gaps of the seeded inputs are >> 1e-5).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def _pair(native_cls, oracle_cls, seed=0):
    torch.manual_seed(seed)
    o = oracle_cls().eval()
    n = native_cls(device="cuda")
    missing = n.load_state_dict(o.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return n, o.to("cuda")


def test_s4_kernel_generation_matches_oracle():
    from oracle import cs3_dgf as O
    from loongx_b200 import cs3

    torch.manual_seed(1)
    for d, n, Ln in [(64, 64, 4096), (4, 4, 256), (6, 6, 512), (6, 6, 128)]:
        lay = cs3.S4Layer(d, n, Ln).to("cuda")
        K = lay.kernel(Ln)
        lam, p, q = O.make_nplr(n)
        # float64 oracle from the float32-stored buffers / parameters (identical inputs to the CUDA kernel)
        c128 = lambda t: t.detach().cpu().to(torch.complex128)  # noqa: E731
        K64 = O.s4_kernel(c128(lay.lambda_), c128(lay.p), c128(lay.q), c128(lay.B), c128(lay.Ct),
                          lay.log_step.detach().cpu().double(), Ln)
        r = _rel(K.cpu(), K64)
        assert r < 1e-5, (d, n, Ln, r)
        assert torch.equal(lay.kernel(Ln), K), "cached kernel must be reused"


@pytest.mark.parametrize("name", ["EEGEncoder", "PPGEncoder", "FNIRSEncoder", "MotionEncoder"])
def test_encoders_match_oracle(name):
    from oracle import cs3_dgf as O
    from loongx_b200 import cs3

    n, o = _pair(getattr(cs3, name), getattr(O, name), seed=3)
    shapes = {"EEGEncoder": (2, 4, 4096), "PPGEncoder": (2, 4, 256), "FNIRSEncoder": (3, 6, 512), "MotionEncoder": (1, 6, 128)}
    g = torch.Generator(device="cuda").manual_seed(45)
    x = torch.randn(shapes[name], generator=g, device="cuda")
    with torch.no_grad():
        ref = o(x)
    got = n(x)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    r = _rel(got, ref)
    print(f"\n[{name}] relL2 {r:.3g}")
    assert r < 1e-4, (name, r)


def test_s4model_reference_layout():
    from oracle import cs3_dgf as O
    from loongx_b200 import cs3

    torch.manual_seed(5)
    o = O.S4Model(4, 16, 8, 2, 16, 300).eval()
    n = cs3.S4Model(4, 16, 8, 2, 16, 300)
    n.load_state_dict(o.state_dict(), strict=True)
    n, o = n.to("cuda"), o.to("cuda")
    u = torch.randn(2, 300, 4, device="cuda")
    with torch.no_grad():
        ref = o(u)
    got = n(u)
    assert _rel(got, ref) < 1e-4


def test_pad_truncate_bit_exact():
    from oracle import cs3_dgf as O
    from loongx_b200 import cs3

    x = torch.randn(2, 4, 5000, device="cuda")
    assert torch.equal(cs3.pad_truncate(x, 4096), O.spatial_pyramid_pooling(x, 4096))
    y = torch.randn(3, 6, 100, device="cuda")
    assert torch.equal(cs3.pad_truncate(y, 128), O.spatial_pyramid_pooling(y, 128))
    z = torch.randn(1, 4, 256, device="cuda")
    assert cs3.pad_truncate(z, 256) is z or torch.equal(cs3.pad_truncate(z, 256), z)


@pytest.mark.parametrize("C,Ln,B", [(512, 4096, 2), (1, 768, 3), (8, 100, 1)])
def test_duan_matches_oracle(C, Ln, B):
    from oracle import cs3_dgf as O
    from loongx_b200 import cs3

    torch.manual_seed(7)
    o = O.DUAN(C).eval()
    n = cs3.DUAN(C, device="cuda")
    n.load_state_dict(o.state_dict(), strict=True)
    o = o.to("cuda")
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn(B, C, Ln, generator=g, device="cuda") * 0.7 + 0.1
    c = torch.randn(B, C, Ln, generator=g, device="cuda")
    with torch.no_grad():
        ref, imp, mask = o(x, c, return_aux=True)
    got = n(x, c)
    torch.cuda.synchronize()
    got_mask = (got.abs().sum(-1) != 0).float()
    k = max(1, int(C * 0.7))
    assert int(mask.sum(1)[0]) == k
    assert torch.equal(got_mask, mask), "top-k channel mask differs from the oracle"
    r = _rel(got, ref)
    print(f"\n[DUAN C={C}] relL2 {r:.3g}, kept {k}/{C}")
    assert r < 1e-4
    # bf16 in -> bf16 out (model.py:995, 1035)
    got16 = n(x.bfloat16(), c.bfloat16())
    assert got16.dtype == torch.bfloat16 and got16.shape == x.shape


def _conditioner_pair():
    from oracle import cs3_dgf as O
    from loongx_b200.config import FluxConfig
    from src.train.model import OminiModel

    torch.manual_seed(11)
    o = O.NeuralConditioner().eval()
    cfg = FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=4096,
                     pooled_projection_dim=768)
    m = OminiModel(cfg, lora_config={"r": 4, "lora_alpha": 4}, device="cuda")
    res = m.load_state_dict(o.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return m, o.to("cuda")


def test_fusion_and_conditioning_modes_match_oracle():
    m, o = _conditioner_pair()
    g = torch.Generator(device="cuda").manual_seed(45)
    B = 2
    eeg = torch.randn(B, 4, 5000, generator=g, device="cuda")   # truncate path
    fn = torch.randn(B, 6, 600, generator=g, device="cuda")
    ppg = torch.randn(B, 4, 256, generator=g, device="cuda")    # exact length
    mo = torch.randn(B, 6, 100, generator=g, device="cuda")     # zero-pad path
    with torch.no_grad():
        pe_ref, po_ref = o.brain_embeddings(eeg, fn, ppg, mo)
    e = m.eeg_projection(m.spatial_pyramid_pooling(eeg, m.eeg_fixed_length))
    p = m.ppg_projection(m.spatial_pyramid_pooling(ppg, m.ppg_fixed_length))
    f = m.fnirs_projection(m.spatial_pyramid_pooling(fn, m.fnirs_fixed_length))
    mm = m.motion_projection(m.spatial_pyramid_pooling(mo, m.motion_fixed_length))
    pe = m.fuse_eeg(e, p)
    po = m.fuse_fnirs(f, mm)
    torch.cuda.synchronize()
    assert _rel(pe, pe_ref) < 1e-4, _rel(pe, pe_ref)
    assert _rel(po, po_ref) < 1e-4, _rel(po, po_ref)
    # generate()-literal fuse: DUAN(x = text, c = brain)
    txt = torch.randn(B, 512, 4096, generator=g, device="cuda") * 0.1
    pooled = torch.randn(B, 768, generator=g, device="cuda")
    with torch.no_grad():
        a_ref, b_ref = o.conditioning(txt, pooled, eeg, fn, ppg, mo, fuse_flag=True, mode="generate")
    a = m.duan_norm_prompt(txt, pe)
    b = m.duan_norm_pooled(pooled.unsqueeze(1), po.unsqueeze(1)).squeeze(1)
    assert _rel(a, a_ref) < 1e-4 and _rel(b, b_ref) < 1e-4
