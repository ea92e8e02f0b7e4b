"""GPU parity tests of the CS3 / DGF backward kernels (csrc/cs3_dgf_bwd.cu, loongx_b200/cs3_bwd.py) against torch
autograd over the oracle's restatement (oracle/cs3_dgf.py), whose gradients equal those the reference's own step() +
backward() leaves on every encoder / fusion parameter (tests/test_reference_pins_cpu.py::
test_reference_encoder_gradients_match_the_oracle, 1e-5).

Stated tolerances (relative L2 per parameter tensor): 1e-4 for the fp32 layers (Linear / LayerNorm / DUAN / pooling /
convolution); 3e-3 for the S4 generating-function parameters (B, Ct, log_step), where the oracle itself evaluates the
Cauchy sums in complex64 while the kernels use float64.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _strict_fp32_oracle():
    """The oracle's nn.Conv1d / matmuls must be true fp32 on the GPU (cuDNN allows TF32 convolutions by default)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _rel(a, b):
    a, b = a.float().flatten(), b.float().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _views(native_module):
    from loongx_b200 import cs3_bwd as CB

    params = list(native_module.parameters())
    flat = torch.zeros(CB.grad_elements(params), device=DEV, dtype=torch.float32)
    return flat, CB.grad_views(params, flat)


def _compare(native_module, views, oracle_module, oracle_grads, tol=1e-4, tol_s4=3e-3, skip=()):
    """oracle_grads: dict name -> grad (None = unused).  Every parameter the oracle differentiates must match."""
    onames = dict(oracle_module.named_parameters())
    worst = {}
    for name, p in native_module.named_parameters():
        if name in skip or name not in onames or oracle_grads.get(name) is None:
            continue
        g = oracle_grads[name]
        want = torch.view_as_real(g.resolve_conj()).flatten() if g.is_complex() else g.flatten()
        got = views[p]
        assert got.numel() == want.numel(), name
        e = _rel(got, want)
        lim = tol_s4 if name.endswith((".s4.B", ".s4.Ct", ".s4.log_step")) else tol
        worst[name] = (e, lim, float(want.norm()))
    bad = {n: v for n, v in worst.items() if v[0] > v[1] and v[2] > 1e-12}
    assert not bad, f"gradient mismatches (relL2, limit, |ref|): {bad}"
    return max(v[0] for v in worst.values()), len(worst)


def _ograds(module, loss):
    names = [n for n, _ in module.named_parameters()]
    gs = torch.autograd.grad(loss, [p for _, p in module.named_parameters()], allow_unused=True)
    return dict(zip(names, gs))


def test_sgemm_ex_variants():
    from loongx_b200 import cs3_bwd as CB

    g = torch.Generator(device=DEV).manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g, device=DEV)  # noqa: E731
    M, N, K, B = 70, 130, 45, 3
    A, Bm = r(B, M, K), r(B, K, N)
    C0 = r(B, M, N)
    for ta in (0, 1):
        for tb in (0, 1):
            As = A.transpose(1, 2).contiguous() if ta else A
            Bs = Bm.transpose(1, 2).contiguous() if tb else Bm
            out = C0.clone()
            CB.sgemm_ex(As, Bs, out, M, N, K, lda=As.shape[2], ldb=Bs.shape[2], ldc=N, trans_a=ta, trans_b=tb, batch=B,
                        a_bs=M * K, b_bs=K * N, c_bs=M * N, alpha=0.5, beta=2.0)
            assert _rel(out, 0.5 * A @ Bm + 2.0 * C0) < 1e-5, (ta, tb)
            red = C0[0].clone()
            CB.sgemm_ex(As, Bs, red, M, N, K, lda=As.shape[2], ldb=Bs.shape[2], ldc=N, trans_a=ta, trans_b=tb, batch=B,
                        a_bs=M * K, b_bs=K * N, reduce=1, beta=1.0)
            assert _rel(red, (A @ Bm).sum(0) + C0[0]) < 1e-5, (ta, tb)


def test_dropout_mask_is_reproducible_and_scaled():
    from loongx_b200 import cs3_bwd as CB

    x = torch.ones(1 << 20, device=DEV)
    a, b, c = CB.dropout(x, 0.3, 7), CB.dropout(x, 0.3, 7), CB.dropout(x, 0.3, 8)
    assert torch.equal(a, b) and not torch.equal(a, c)
    kept = (a != 0).float().mean().item()
    assert abs(kept - 0.7) < 3e-3 and abs(a.mean().item() - 1.0) < 5e-3
    assert set(a.unique().tolist()) == {0.0, float(torch.tensor(1.0 / 0.7, dtype=torch.float32))}


@pytest.mark.parametrize("tokens", [True, False])
def test_projection_mlp_backward(tokens):
    from oracle import cs3_dgf as OC
    from loongx_b200 import cs3, cs3_bwd as CB

    torch.manual_seed(3)
    d_in, d_hid = 360, 256
    om = (OC._mlp_tokens(d_in, d_hid) if tokens else OC._mlp_pooled(d_in, d_hid)).to(DEV)
    nm = cs3._projection(d_in, d_hid, 4096 if tokens else 768, tokens).to(DEV)
    nm.load_state_dict(om.state_dict())
    B = 3
    feat = torch.randn(B, d_in, device=DEV)
    out_o = om(feat)
    gup = torch.randn_like(out_o)
    og = _ograds(om, (out_o * gup).sum())
    out_n, ctx = CB.projection_forward_train(nm, feat.clone(), training=False, seed=0)
    assert _rel(out_n, out_o) < 1e-5
    flat, views = _views(nm)
    dfeat = CB.projection_backward(nm, ctx, gup, views)
    featg = feat.clone().requires_grad_(True)
    want_dfeat, = torch.autograd.grad((om(featg) * gup).sum(), featg)
    worst, n = _compare(nm, views, om, og)
    print(f"\n[projection bwd tokens={tokens}] {n} tensors, worst relL2 {worst:.3g}, d feat {_rel(dfeat, want_dfeat):.3g}")
    assert _rel(dfeat, want_dfeat) < 1e-4


def test_projection_mlp_dropout_train_mode():
    """Train mode: the forward output equals the eval pipeline with the two masks applied, and the backward uses the same
    masks (checked through a finite difference on one weight)."""
    from loongx_b200 import cs3, cs3_bwd as CB

    torch.manual_seed(4)
    nm = cs3._projection(200, 128, 768, False).to(DEV)
    feat = torch.randn(2, 200, device=DEV)
    o1, c1 = CB.projection_forward_train(nm, feat, training=True, seed=5)
    o2, _ = CB.projection_forward_train(nm, feat, training=True, seed=5)
    o3, _ = CB.projection_forward_train(nm, feat, training=True, seed=6)
    assert torch.equal(o1, o2) and not torch.equal(o1, o3)
    assert 0.2 < (o1 == 0).float().mean().item() < 0.9  # ReLU zeros + dropped units
    gup = torch.randn_like(o1)
    flat, views = _views(nm)
    CB.projection_backward(nm, c1, gup, views)
    w = nm[5].weight
    i, j = 3, 7
    eps = 1e-2
    with torch.no_grad():
        w[i, j] += eps
        lp = (CB.projection_forward_train(nm, feat, True, 5)[0] * gup).sum().item()
        w[i, j] -= 2 * eps
        lm = (CB.projection_forward_train(nm, feat, True, 5)[0] * gup).sum().item()
        w[i, j] += eps
    fd = (lp - lm) / (2 * eps)
    got = views[w].view_as(w)[i, j].item()
    assert abs(fd - got) <= 2e-2 * max(1.0, abs(fd)), (fd, got)


@pytest.mark.parametrize("d,n,L,B", [(4, 4, 256, 2), (6, 6, 128, 3), (16, 16, 300, 2), (64, 64, 1024, 1)])
def test_s4model_backward(d, n, L, B):
    from oracle import cs3_dgf as OC
    from loongx_b200 import cs3, cs3_bwd as CB

    torch.manual_seed(5)
    om = OC.S4Model(d if d < 64 else 4, d, d, 2, n, L).to(DEV)
    nm = cs3.S4Model(d if d < 64 else 4, d, d, 2, n, L).to(DEV)
    nm.load_state_dict(om.state_dict())
    d_in = om.encoder.in_features
    x = torch.randn(B, d_in, L, device=DEV)
    out_o = om(x.permute(0, 2, 1)).permute(0, 2, 1)  # channel-major [B, d, L]
    gup = torch.randn_like(out_o)
    og = _ograds(om, (out_o * gup).sum())
    out_n, ctx = CB.s4model_forward_train(nm, x)
    assert _rel(out_n, out_o) < 1e-4
    flat, views = _views(nm)
    CB.s4model_backward(nm, ctx, gup.contiguous(), views)
    worst, cnt = _compare(nm, views, om, og)
    print(f"\n[S4Model bwd d={d} n={n} L={L}] {cnt} tensors, worst relL2 {worst:.3g}")


@pytest.mark.parametrize("Cc,L,B", [(512, 64, 2), (1, 768, 3), (6, 100, 2)])
def test_duan_backward(Cc, L, B):
    from oracle import cs3_dgf as OC
    from loongx_b200 import cs3, cs3_bwd as CB

    torch.manual_seed(6)
    om = OC.DUAN(Cc).to(DEV)
    nm = cs3.DUAN(Cc, device=DEV)
    nm.load_state_dict(om.state_dict())
    x = torch.randn(B, Cc, L, device=DEV) * 1.5 + 0.3
    c = torch.randn(B, Cc, L, device=DEV)
    xg, cg = x.clone().requires_grad_(True), c.clone().requires_grad_(True)
    out_o = om(xg, cg)
    gup = torch.randn_like(out_o)
    loss = (out_o * gup).sum()
    dx_o, dc_o = torch.autograd.grad(loss, (xg, cg), retain_graph=True)
    og = _ograds(om, loss)
    out_n, ctx = CB.duan_forward_train(nm, x, c)
    assert _rel(out_n, out_o) < 1e-4
    flat, views = _views(nm)
    dx, dc = torch.empty_like(x), torch.empty_like(c)
    CB.duan_backward(nm, ctx, gup.contiguous(), views, dx=dx, dc=dc)
    worst, cnt = _compare(nm, views, om, og)
    print(f"\n[DUAN bwd C={Cc} L={L}] {cnt} tensors, worst relL2 {worst:.3g}; dx {_rel(dx, dx_o):.3g} dc {_rel(dc, dc_o):.3g}")
    assert _rel(dx, dx_o) < 1e-4 and _rel(dc, dc_o) < 1e-4


@pytest.mark.parametrize("which", ["ppg", "fnirs", "motion", "eeg"])
def test_encoder_backward(which):
    from oracle import cs3_dgf as OC
    from loongx_b200 import cs3, cs3_bwd as CB

    torch.manual_seed(7)
    pairs = dict(ppg=(OC.PPGEncoder, cs3.PPGEncoder, (4, 256)), fnirs=(OC.FNIRSEncoder, cs3.FNIRSEncoder, (6, 512)),
                 motion=(OC.MotionEncoder, cs3.MotionEncoder, (6, 128)), eeg=(OC.EEGEncoder, cs3.EEGEncoder, (4, 4096)))
    Oc, Nc, (ch, L) = pairs[which]
    om = Oc().to(DEV)
    nm = Nc(device=DEV)
    nm.load_state_dict(om.state_dict())
    B = 2
    x = torch.randn(B, ch, L, device=DEV)
    out_o = om(x)
    gup = torch.randn_like(out_o) / out_o.numel() ** 0.5
    og = _ograds(om, (out_o * gup).sum())
    out_n, ctx = CB.encoder_forward_train(nm, x, training=False, seed=0)
    assert _rel(out_n, out_o) < 1e-4
    flat, views = _views(nm)
    CB.encoder_backward(nm, ctx, gup.contiguous(), views)
    worst, cnt = _compare(nm, views, om, og)
    print(f"\n[{which} encoder bwd] {cnt} tensors, worst relL2 {worst:.3g}")


@pytest.mark.parametrize("fuse", [True, False])
def test_step_conditioning_backward_vs_oracle(fuse):
    """The whole conditioning of OminiModel.step (model.py:656-701), all four signals: every CS3 / DGF parameter gradient
    for a random upstream gradient on (prompt_embeds, pooled)."""
    from oracle import cs3_dgf as OC
    from loongx_b200 import cs3_bwd as CB
    from loongx_b200.config import FluxConfig
    from src.train.model import OminiModel

    torch.manual_seed(11)
    oc = OC.NeuralConditioner().to(DEV)
    kw = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=4096, pooled_projection_dim=768)
    m = OminiModel(FluxConfig(**kw), lora_config={"r": 4, "lora_alpha": 4}, device=DEV, model_config={}, fuse_flag=fuse)
    m.load_state_dict(oc.state_dict(), strict=True)
    g = torch.Generator().manual_seed(3)
    B = 2
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(DEV)  # noqa: E731
    pe, po = r(B, 512, 4096, scale=0.3), r(B, 768)
    sig = dict(eeg=r(B, 4, 5000), fnirs=r(B, 6, 600), ppg=r(B, 4, 256), motion=r(B, 6, 100))
    pe_o, po_o = oc.conditioning(pe, po, sig["eeg"], sig["fnirs"], sig["ppg"], sig["motion"], fuse_flag=fuse, mode="step")
    g1, g2 = torch.randn_like(pe_o) / 1e3, torch.randn_like(po_o) / 30
    og = _ograds(oc, (pe_o * g1).sum() + (po_o * g2).sum())
    pe_n, po_n, ctx = CB.step_conditioning_train(m, pe, po, sig["eeg"], sig["fnirs"], sig["ppg"], sig["motion"], training=False,
                                                 seed=0)
    assert _rel(pe_n, pe_o) < 5e-3 and _rel(po_n, po_o) < 5e-3  # outputs are rounded to bf16 for the DiT
    params = CB.trainable_parameters(m)
    flat = torch.zeros(CB.grad_elements(params), device=DEV)
    views = CB.grad_views(params, flat)
    CB.step_conditioning_backward(m, ctx, g1, g2, views)
    used = {n for n, v in og.items() if v is not None}
    worst, cnt = _compare(m, views, oc, og)
    print(f"\n[step conditioning bwd fuse={fuse}] {cnt} tensors compared ({len(used)} used by the oracle), worst relL2 {worst:.3g}")
    assert cnt == len(used)
    # parameters the oracle does not reach must stay untouched
    for n, p in m.named_parameters():
        if og.get(n) is None:
            assert float(views[p].abs().max()) == 0.0, n


def test_dit_input_gradients_vs_oracle():
    """d loss / d prompt_embeds and d loss / d pooled_projections out of the native DiT backward against fp32 autograd of
    the oracle's flow_step (no brain conditioning): the gradients that enter the CS3 / DGF backward."""
    from oracle import flux_dit as O
    from oracle import train_step as TS
    from loongx_b200.config import FluxConfig
    from loongx_b200.dit import DitWeights
    from loongx_b200.train import DitTrainer

    kw = dict(num_layers=2, num_single_layers=2, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
    ocfg, cfg = O.FluxConfig(**kw), FluxConfig(**kw)
    P = O.init_params(ocfg, seed=21, dtype=torch.float32, device="cpu", w_std=0.05, bias_std=0.05, lora_b_std=0.05)
    Pb = {k: v.to(torch.bfloat16).to(DEV) for k, v in P.items()}
    P32 = {k: v.float() for k, v in Pb.items()}
    W = DitWeights(dict(Pb), cfg, DEV)
    g = torch.Generator().manual_seed(8)
    B, h, w = 2, 16, 32
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).bfloat16().to(DEV)  # noqa: E731
    batch = dict(image=r(B, 16, h, w), condition=r(B, 16, h, w), prompt_embeds=r(B, 128, 256, scale=0.5),
                 pooled_prompt_embeds=r(B, 64), position_delta=[[0, -16]], t=torch.tensor([0.3, 0.7]), noise=r(B, 128, 64))
    # oracle
    b32 = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
    b32["prompt_embeds"].requires_grad_(True)
    b32["pooled_prompt_embeds"].requires_grad_(True)
    loss_o, _ = TS.flow_step(P32, ocfg, b32, model_config={})
    d_pe_o, d_po_o = torch.autograd.grad(loss_o, (b32["prompt_embeds"], b32["pooled_prompt_embeds"]))
    # native
    from oracle import sampler as OS

    tr = DitTrainer(W, B, 128, 128, 128, model_config={}, input_grads=True)
    x0 = OS.pack_latents(batch["image"])
    cond = OS.pack_latents(batch["condition"])
    img_ids = OS.prepare_latent_image_ids(h, w).to(DEV)
    cond_ids = OS.condition_ids(OS.prepare_latent_image_ids(h, w).to(DEV), [0, -16], 1.0)
    loss = tr.forward(x0.contiguous(), batch["noise"], batch["t"].to(DEV), cond, batch["prompt_embeds"],
                      batch["pooled_prompt_embeds"], torch.zeros(128, 3, device=DEV), img_ids.float(), cond_ids, guidance=1.0)
    tr.zero_grad()
    tr.backward(1.0)
    e_pe, e_po = _rel(tr.d_prompt, d_pe_o), _rel(tr.d_pooled, d_po_o)
    print(f"\n[DiT input grads] loss {float(loss):.5f} (oracle {float(loss_o):.5f}); relL2 d prompt_embeds {e_pe:.3g}, d pooled {e_po:.3g}")
    assert e_pe < 3e-2 and e_po < 3e-2


def test_step_leaves_encoder_gradients_like_the_reference():
    """OminiModel.step + loss.backward(): `.grad` on every CS3 / DGF parameter the reference's autograd reaches, against
    fp32 autograd over the oracle's flow_step (which equals the reference's own step() + backward(), 1e-5).  The DiT between
    the loss and the conditioning runs in bf16: tolerance 5e-2 relL2 per tensor (and 2e-2 on the concatenation)."""
    from oracle import cs3_dgf as OC
    from oracle import flux_dit as O
    from oracle import train_step as TS
    from loongx_b200.config import FluxConfig
    from src.train.model import OminiModel

    torch.manual_seed(13)
    oc = OC.NeuralConditioner().to(DEV)
    kw = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=4096, pooled_projection_dim=768)
    cfg = FluxConfig(**kw)
    m = OminiModel(cfg, lora_config={"r": 4, "lora_alpha": 4}, device=DEV, model_config={}, use_brain_condition=True,
                   fuse_flag=True, seed=5)
    m.load_state_dict(oc.state_dict(), strict=True)
    g = torch.Generator().manual_seed(4)
    B, h, w = 2, 16, 32
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(DEV)  # noqa: E731
    batch = dict(image=r(B, 16, h, w).bfloat16(), condition=r(B, 16, h, w).bfloat16(),
                 prompt_embeds=r(B, 512, 4096, scale=0.3).bfloat16(), pooled_prompt_embeds=r(B, 768).bfloat16(),
                 position_delta=[[0, -16]], condition_type=["subject"] * B, t=torch.tensor([0.4, 0.8]),
                 noise=r(B, 128, 64).bfloat16(), eeg=r(B, 4, 5000), fnirs=r(B, 6, 600), ppg=r(B, 4, 256), motion=r(B, 6, 100))
    loss = m.step(batch)
    loss.backward()
    tr = m._trainer_obj
    assert tr.grad_flat.numel() > tr.n_lora_grad + 59_000_000  # LoRA + ~59.4 M encoder elements in ONE bucket
    # oracle with the native DiT weights
    P = {k: v.float() for k, v in m.transformer.weights.export_params().items()}
    ocfg = O.FluxConfig(**kw)
    b32 = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
    loss_o, _ = TS.flow_step(P, ocfg, b32, model_config={}, conditioner=oc, use_brain_condition=True, fuse_flag=True)
    og = _ograds(oc, loss_o)
    got_all, want_all, n = [], [], 0
    worst = 0.0
    for name, p in m.named_parameters():
        if og.get(name) is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        a = torch.view_as_real(p.grad.resolve_conj()).flatten() if p.grad.is_complex() else p.grad.flatten()
        b_ = torch.view_as_real(og[name].resolve_conj()).flatten() if og[name].is_complex() else og[name].flatten()
        if float(b_.norm()) > 1e-10:
            worst = max(worst, _rel(a, b_))
        got_all.append(a)
        want_all.append(b_)
        n += 1
    tot = _rel(torch.cat(got_all), torch.cat(want_all))
    # per tensor: those that carry a non-negligible share of the gradient (tiny ones are bf16 noise of the DiT in between)
    big = max(float(b_.norm()) for b_ in want_all)
    worst = max(_rel(a, b_) for a, b_ in zip(got_all, want_all) if float(b_.norm()) > 1e-2 * big)
    print(f"\n[step encoder grads] loss {float(loss):.5f} (oracle {float(loss_o):.5f}); {n} tensors, worst relL2 {worst:.3g}, "
          f"all parameters together {tot:.3g}")
    assert abs(float(loss) - float(loss_o)) / float(loss_o) < 1e-2
    assert tot < 2e-2 and worst < 5e-2
    assert any(p.grad is not None for p in m.lora_layers)
