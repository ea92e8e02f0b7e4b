"""GPU parity of the training step (a17: OminiModel.step, model.py:569-729): every backward / un-fused forward kernel
against torch autograd of the same fp32 formula, then the whole native forward + backward (loss and LoRA-factor
gradients) against the oracle restatement of the reference's step — which tests/test_reference_pins_cpu.py pins
bit-for-bit to the reference's own `step()` + `loss.backward()`.

Tolerances (bf16 storage of activations and activation gradients, fp32 accumulation): kernels relL2 <= 1e-2 against
fp32 autograd on the same bf16-rounded inputs; end to end loss rel <= 2e-2, concatenated LoRA gradients relL2 <= 6e-2.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _geo(B=2, nt=128, ni=256, nc=128):
    from loongx_b200 import ops

    tm = ops.make_tile_meta(B, nt, ni, nc, DEV)
    R = B * (nt + ni + nc)
    return tm, R


def _stream_batch_of_rows(tm, R):
    t = tm.cpu()
    stream = t[:, 0].repeat_interleave(128)[:R].to(DEV)
    batch = t[:, 1].repeat_interleave(128)[:R].to(DEV)
    return stream.long(), batch.long()


def _rand(*s, scale=1.0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (torch.randn(*s, generator=g, device=DEV) * scale).bfloat16()


def test_gelu_fwd_bwd():
    from loongx_b200 import train as T

    x = _rand(256, 1024, scale=2.0)
    dy = _rand(256, 1024, seed=1)
    out = torch.empty_like(x)
    T.gelu_fwd(x, out)
    xf = x.float().requires_grad_(True)
    ref = F.gelu(xf, approximate="tanh")
    assert _rel(out, ref) < 5e-3
    (gref,) = torch.autograd.grad(ref, xf, dy.float())
    dx = torch.empty_like(x)
    T.gelu_bwd(x, dy, dx)
    assert _rel(dx, gref) < 5e-3
    big = torch.zeros(256, 2048, device=DEV, dtype=torch.bfloat16)  # strided views + in place
    big[:, 512:1536] = dy
    T.gelu_bwd(x, big[:, 512:1536], big[:, 512:1536])
    assert torch.equal(big[:, 512:1536], dx)


def test_gate_residual_and_gate_bwd():
    from loongx_b200 import train as T

    B, D = 2, 512
    tm, R = _geo(B)
    stream, batch = _stream_batch_of_rows(tm, R)
    res, y, dout = _rand(R, D), _rand(R, D, seed=1), _rand(R, D, seed=2)
    gates = [_rand(B, 3 * D, seed=3 + s)[:, D:2 * D] for s in range(3)]  # strided [B, D] views
    gfull = torch.stack([g.float() for g in gates])[stream, batch]  # [R, D]
    out = torch.empty_like(res)
    T.gate_residual_fwd(res, y, out, tm, gates)
    assert _rel(out, res.float() + gfull * y.float()) < 4e-3
    dy = torch.empty_like(res)
    dg = torch.zeros(B, 2 * D, device=DEV)
    T.gate_bwd(dout, y, dy, tm, gates, [None, None, dg[:, D:]])
    assert _rel(dy, gfull * dout.float()) < 4e-3
    prod = dout.float() * y.float()
    ref = torch.stack([prod[(stream == 2) & (batch == b)].sum(0) for b in range(B)])
    assert _rel(dg[:, D:], ref) < 1e-4 and float(dg[:, :D].abs().max()) == 0.0


@pytest.mark.parametrize("D", [256, 3072])
def test_ln_modulate_bwd(D):
    from loongx_b200 import train as T

    B = 2
    tm, R = _geo(B)
    stream, batch = _stream_batch_of_rows(tm, R)
    x, dxn, dres = _rand(R, D, scale=1.5), _rand(R, D, seed=1), _rand(R, D, seed=2)
    scales = [_rand(B, D, scale=0.3, seed=3 + s) for s in range(3)]
    xf = x.float().requires_grad_(True)
    sc = torch.stack([s.float() for s in scales]).requires_grad_(True)  # [3, B, D]
    sh = torch.zeros_like(sc).requires_grad_(True)
    xn = F.layer_norm(xf, (D,), eps=1e-6) * (1 + sc[stream, batch]) + sh[stream, batch]
    gx, gsc, gsh = torch.autograd.grad(xn, (xf, sc, sh), dxn.float())
    dx = torch.empty_like(x)
    dscale = [torch.zeros(B, D, device=DEV) for _ in range(3)]
    dshift = [torch.zeros(B, D, device=DEV) for _ in range(3)]
    stats = torch.zeros(R, 2, device=DEV)
    T.ln_modulate_bwd(x, dxn, dres, dx, tm, scales, dscale, dshift, stats)
    assert _rel(dx, gx + dres.float()) < 6e-3
    for s in range(3):
        assert _rel(dscale[s], gsc[s]) < 2e-3 and _rel(dshift[s], gsh[s]) < 1e-4
    # no residual, in place into dxn's buffer is not allowed but dres == dx is
    dx2 = dres.clone()
    T.ln_modulate_bwd(x, dxn, dx2, dx2, tm, scales, [None] * 3, [None] * 3, None)
    assert torch.equal(dx2, dx)


def test_qkv_post_fwd_bwd_and_rows_to_heads():
    from oracle import flux_dit as O
    from loongx_b200 import ops
    from loongx_b200 import train as T

    B, H, nt, ni, nc = 2, 2, 128, 128, 128
    S = nt + ni + nc
    tm = ops.make_tile_meta(B, nt, ni, nc, DEV)
    orb = ops.make_out_row_base(B, nt, ni, nc, DEV)
    R, D = B * S, H * 128
    pre = torch.zeros(R, 7 * D, device=DEV, dtype=torch.bfloat16)
    pre[:, :3 * D] = _rand(R, 3 * D, scale=1.3)
    wq = [(1 + 0.1 * torch.randn(128, device=DEV)).float() for _ in range(2)]
    wk = [(1 + 0.1 * torch.randn(128, device=DEV)).float() for _ in range(2)]
    ids = torch.cat([torch.zeros(nt, 3), torch.rand(ni + nc, 3) * 20], 0).to(DEV)
    ids[:, 0] = 0
    rope = torch.zeros(S, 64, 2, device=DEV)
    from loongx_b200.dit import _lib, _stream
    from loongx_b200 import _lib as L

    L.check(_lib.lx_rope_table(ids.contiguous().data_ptr(), rope.data_ptr(), S, 16, 56, 56, 10000.0, _stream()), "rope")
    q, k, v = (torch.zeros(B, H, S, 128, device=DEV, dtype=torch.bfloat16) for _ in range(3))
    T.qkv_post_fwd(pre, H, tm, q, k, v, [wq[0], wq[1], wq[1]], [wk[0], wk[1], wk[1]], rope)

    # reference: rows -> [B, S, 3D] in joint-sequence order
    def to_seq(rows):  # [R, C] -> [B, S, C]
        rt, ri = B * nt, B * ni
        return torch.cat([rows[:rt].view(B, nt, -1), rows[rt:rt + ri].view(B, ni, -1), rows[rt + ri:].view(B, nc, -1)], 1)

    def from_seq(x):  # [B, S, C] -> [R, C]
        return torch.cat([x[:, :nt].reshape(B * nt, -1), x[:, nt:nt + ni].reshape(B * ni, -1), x[:, nt + ni:].reshape(B * nc, -1)])

    cos, sin = O.rope_tables(ids)
    x = to_seq(pre[:, :3 * D].float()).requires_grad_(True)

    def ref_fwd(x):
        outs = []
        for which in range(3):
            t = x[..., which * D:(which + 1) * D].view(B, S, H, 128).transpose(1, 2)  # [B,H,S,128]
            if which < 2:
                w = wq if which == 0 else wk
                wt = torch.cat([w[0].expand(nt, 128), w[1].expand(ni + nc, 128)], 0)  # per-position weight
                var = t.pow(2).mean(-1, keepdim=True)
                t = t * torch.rsqrt(var + 1e-6) * wt[None, None]
                t = O.apply_rotary_emb(t, (cos, sin))
            outs.append(t)
        return outs

    rq, rk, rv = ref_fwd(x)
    for got, ref in ((q, rq), (k, rk), (v, rv)):
        assert _rel(got, ref) < 5e-3
    dq, dk, dv = _rand(B, H, S, 128, seed=5), _rand(B, H, S, 128, seed=6), _rand(B, H, S, 128, seed=7)
    (gx,) = torch.autograd.grad([rq, rk, rv], x, [dq.float(), dk.float(), dv.float()])
    dpre = torch.zeros(R, 7 * D, device=DEV, dtype=torch.bfloat16)
    T.qkv_post_bwd(pre, dq, dk, dv, dpre, H, tm, [wq[0], wq[1], wq[1]], [wk[0], wk[1], wk[1]], rope)
    assert _rel(dpre[:, :3 * D], from_seq(gx)) < 6e-3
    assert float(dpre[:, 3 * D:].abs().max()) == 0.0
    # rows_to_heads is the inverse of the attention kernel's output layout
    rows = _rand(R, 5 * D, seed=9)
    heads = torch.zeros(B, H, S, 128, device=DEV, dtype=torch.bfloat16)
    T.rows_to_heads(rows, H, tm, heads)
    ref_heads = to_seq(rows[:, :D]).view(B, S, H, 128).transpose(1, 2)
    assert torch.equal(heads, ref_heads)
    assert orb.numel() == B * S // 128


@pytest.mark.parametrize("M,K,N,r", [(256, 256, 512, 4), (384, 64, 256, 4), (2, 256, 1536, 4), (130, 1280, 256, 8)])
def test_lora_grad_merge_transpose(M, K, N, r):
    from loongx_b200 import train as T

    x, dy = _rand(M, K), _rand(M, N, seed=1)
    A = torch.randn(r, K, device=DEV) / r
    Bw = torch.randn(N, r, device=DEV) * 0.05
    s = 1.7
    Ag, Bg = A.clone().requires_grad_(True), Bw.clone().requires_grad_(True)
    y = (x.float() @ Ag.t()) @ Bg.t() * s
    gA, gB = torch.autograd.grad(y, (Ag, Bg), dy.float())
    dA, dB = torch.zeros_like(A), torch.zeros_like(Bw)
    ws = torch.zeros(2 * M * r, device=DEV)
    T.lora_grad(x, dy, A, Bw, dA, dB, s, ws)
    assert _rel(dA, gA) < 1e-4 and _rel(dB, gB) < 1e-4
    T.lora_grad(x, dy, A, Bw, dA, dB, s, ws)  # accumulates
    assert _rel(dA, 2 * gA) < 1e-4
    W = _rand(N, K, scale=0.05, seed=2)
    out = torch.empty_like(W)
    T.lora_merge(W, A, Bw, out, s)
    assert torch.equal(out, torch.addmm(W.float(), Bw, A, alpha=s).bfloat16()) or _rel(out, W.float() + s * Bw @ A) < 3e-3
    assert torch.equal(T.transpose(W), W.t().contiguous())
    if N % 8 == 0 and K % 8 == 0:  # both panels in one pass (what LoraFactor.remerge uses): bit-identical to merge + transpose
        big, bigT = torch.zeros((N + 16, K), device=DEV, dtype=torch.bfloat16), torch.zeros((K, N + 16), device=DEV, dtype=torch.bfloat16)
        T.lora_merge_t(W, A, Bw, big[8:8 + N], bigT[:, 8:8 + N], s)
        assert torch.equal(big[8:8 + N], out) and torch.equal(bigT[:, 8:8 + N], out.t())
        assert (big[:8] == 0).all() and (big[8 + N:] == 0).all() and (bigT[:, :8] == 0).all() and (bigT[:, 8 + N:] == 0).all()


@pytest.mark.parametrize("M,K,widths,r", [(4096 + 40, 768, (256, 256, 256), 4), (300, 256, (512, 256, 256, 1024), 4),
                                          (64, 64, (768,), 4), (130, 512, (256, 512), 8), (257, 256, (512,), 16),
                                          (1000, 1024, (2048,), 4), (5, 256, (512, 256), 4), (1, 512, (256,), 4),
                                          (4, 3072, (1024, 1024, 1024, 1024), 4)])
def test_lora_grad_stacked_vs_autograd_and_per_factor(M, K, widths, r):
    """lx_lora_grad_stacked: the sub-Linears of one fused projection (to_q | to_k | to_v (| proj_mlp)) in four launches;
    fp32 autograd over y_g = s_g (x A_g^T) B_g^T is the checker, and the per-factor kernels must agree."""
    from types import SimpleNamespace as NS

    from loongx_b200 import train as T

    N = sum(widths)
    pad = 24  # a row stride wider than the data, like the column views the trainer hands in
    xb, dyb = _rand(M, K + pad), _rand(M, N + pad, seed=1)
    x, dy = xb[:, :K], dyb[:, :N]
    fs, refs, c = [], [], 0
    for g, w in enumerate(widths):
        A = torch.randn(r, K, device=DEV) / r
        Bw = torch.randn(w, r, device=DEV) * 0.05
        s = 0.6 + 0.4 * g
        Ag, Bg = A.clone().requires_grad_(True), Bw.clone().requires_grad_(True)
        y = (x.float() @ Ag.t()) @ Bg.t() * s
        refs.append(torch.autograd.grad(y, (Ag, Bg), dy[:, c:c + w].float()))
        fs.append(NS(A=A, B=Bw, dA=torch.zeros_like(A), dB=torch.zeros_like(Bw), rows=w, panel=NS(scaling=s)))
        c += w
    assert T.lora_grad_stackable(fs)
    ws = torch.zeros(2 * M * 16, device=DEV)
    T.lora_grad_stacked(x, dy, fs, ws)
    c = 0
    for f, (gA, gB) in zip(fs, refs):
        assert _rel(f.dA, gA) < 1e-4 and _rel(f.dB, gB) < 1e-4
        dA, dB = torch.zeros_like(f.A), torch.zeros_like(f.B)
        T.lora_grad(x, dy[:, c:c + f.rows], f.A, f.B, dA, dB, f.panel.scaling, ws)
        assert _rel(f.dA, dA) < 1e-5 and _rel(f.dB, dB) < 1e-5
        c += f.rows
    T.lora_grad_stacked(x, dy, fs, ws)  # accumulates
    assert _rel(fs[0].dA, 2 * refs[0][0]) < 1e-4 and _rel(fs[-1].dB, 2 * refs[-1][1]) < 1e-4


def test_flow_objective_kernels():
    from loongx_b200 import train as T

    B, n, C = 3, 128, 64
    x0, x1, pred = _rand(B, n, C), _rand(B, n, C, seed=1), _rand(B, n, C, seed=2)
    t = torch.tensor([0.1, 0.5, 0.93], device=DEV)
    xt = T.flow_noise_mix(x0, x1, t)
    ref = ((1 - t[:, None, None]) * x0.float() + t[:, None, None] * x1.float())
    assert _rel(xt, ref) < 4e-3
    loss = torch.zeros(1, device=DEV)
    dpred = torch.empty_like(pred)
    T.flow_mse_loss(pred, x0, x1, loss, dpred, 1.0)
    pf = pred.float().requires_grad_(True)
    lref = F.mse_loss(pf, (x1 - x0).float())
    (gref,) = torch.autograd.grad(lref, pf)
    assert abs(loss.item() - lref.item()) / lref.item() < 1e-5
    assert _rel(dpred, gref) < 4e-3


def _tiny_train_case(layers=(2, 2), B=2, nt=128, hw=(16, 32), seed=0):
    from oracle import flux_dit as O
    from loongx_b200.config import FluxConfig

    kw = dict(num_layers=layers[0], num_single_layers=layers[1], num_attention_heads=2, joint_attention_dim=256,
              pooled_projection_dim=64)
    ocfg, cfg = O.FluxConfig(**kw), FluxConfig(**kw)
    P = O.init_params(ocfg, seed=1234, dtype=torch.float32, device="cpu", w_std=0.05, bias_std=0.05, lora_b_std=0.05)
    P = {k: v.to(torch.bfloat16) for k, v in P.items()}
    g = torch.Generator().manual_seed(seed)
    h, w = hw  # latent grid: (h/2)*(w/2) image tokens
    ni = (h // 2) * (w // 2)
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).bfloat16()  # noqa: E731
    batch = dict(image=r(B, 16, h, w), condition=r(B, 16, h, w), prompt_embeds=r(B, nt, 256, scale=0.5),
                 pooled_prompt_embeds=r(B, 64), position_delta=[[0, -(w // 2)]], t=torch.tensor([0.35, 0.8][:B]),
                 noise=r(B, ni, 64))
    return ocfg, cfg, P, batch


@pytest.mark.parametrize("recompute", [False, True])
@pytest.mark.parametrize("layers,mc,nt,hw", [((1, 1), {}, 128, (16, 32)), ((2, 2), {}, 128, (16, 32)),
                                             ((2, 2), {"latent_lora": True}, 128, (16, 32)),
                                             ((1, 2), {"independent_condition": True}, 128, (16, 32)),
                                             ((2, 1), {"add_cond_attn": True}, 128, (16, 32)),
                                             ((1, 1), {}, 100, (12, 20)), ((2, 1), {"latent_lora": True}, 77, (20, 36))])
def test_train_step_loss_and_lora_grads_vs_oracle(layers, mc, nt, hw, recompute):
    """Native forward + backward vs fp32 autograd over the oracle restatement of model.py:569-729."""
    from oracle import sampler as OS
    from oracle import train_step as TS
    from loongx_b200.dit import DitWeights
    from loongx_b200.train import DitTrainer

    ocfg, cfg, P, batch = _tiny_train_case(layers, nt=nt, hw=hw)
    B = batch["image"].shape[0]
    ni = (hw[0] // 2) * (hw[1] // 2)
    P32 = {k: v.float().to(DEV) for k, v in P.items()}
    b_dev = {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    loss_ref, grads_ref, aux = TS.flow_step_grads(P32, ocfg, {k: (v.float() if isinstance(v, torch.Tensor) else v)
                                                              for k, v in b_dev.items()}, model_config=mc)
    W = DitWeights({k: v.to(DEV) for k, v in P.items()}, cfg, DEV)
    tr = DitTrainer(W, B, nt, ni, ni, model_config=mc, recompute=recompute)  # ragged lengths are padded inside
    assert tr.recompute == recompute
    x0 = OS.pack_latents(b_dev["image"]).contiguous()
    cond = OS.pack_latents(b_dev["condition"]).contiguous()
    img_ids = OS.prepare_latent_image_ids(batch["image"].shape[2], batch["image"].shape[3]).to(DEV)
    cond_ids = OS.condition_ids(img_ids, batch["position_delta"][0])
    loss = tr.forward(x0, b_dev["noise"], b_dev["t"], cond, b_dev["prompt_embeds"], b_dev["pooled_prompt_embeds"],
                      torch.zeros(nt, 3, device=DEV), img_ids, cond_ids, guidance=1.0)
    torch.cuda.synchronize()
    loss1 = loss.item()  # the trainer reuses its loss buffer
    e_pred = _rel(tr.pred, aux["pred"])
    print(f"\n[train {layers} {mc} nt={nt} ni={ni} recompute={recompute}] loss native {loss1:.6f} oracle {loss_ref.item():.6f}  pred relL2 {e_pred:.4g}")
    assert e_pred < 2e-2
    assert abs(loss1 - loss_ref.item()) / loss_ref.item() < 2e-2
    tr.zero_grad()
    tr.backward()
    torch.cuda.synchronize()
    got = tr.grads()
    assert set(got) == set(grads_ref)
    names = sorted(got)
    cat_g = torch.cat([got[n].flatten() for n in names])
    cat_r = torch.cat([grads_ref[n].flatten() for n in names])
    e_all = _rel(cat_g, cat_r)
    worst = max(((_rel(got[n], grads_ref[n]), n) for n in names if grads_ref[n].norm() > 1e-3 * cat_r.norm()), default=(0, ""))
    print(f"[train {layers}] LoRA grads relL2 {e_all:.4g}; worst significant param {worst[1]} {worst[0]:.4g}")
    assert torch.isfinite(cat_g).all()
    assert e_all < 6e-2, e_all
    assert worst[0] < 0.15, worst
    for n in names:  # parameters the loss does not depend on (last block's condition-only paths) get exactly zero
        if float(grads_ref[n].abs().max()) == 0.0:
            assert float(got[n].abs().max()) < 1e-6 * float(cat_r.abs().max()) + 1e-12, n
    # an SGD step on the factors followed by remerge changes the loss in the descent direction
    lr = 0.5 / float(cat_g.norm())
    for f in tr.factors.values():
        f.A.data.add_(f.dA, alpha=-lr)
        f.B.data.add_(f.dB, alpha=-lr)
    tr.remerge()
    loss2 = tr.forward(x0, b_dev["noise"], b_dev["t"], cond, b_dev["prompt_embeds"], b_dev["pooled_prompt_embeds"],
                       torch.zeros(nt, 3, device=DEV), img_ids, cond_ids, guidance=1.0).item()
    print(f"[train {layers}] loss after one SGD step {loss2:.6f}")
    assert loss2 < loss1


def test_omini_model_step_api_and_optimizer():
    """OminiModel.step (model.py:569-729) through the drop-in module surface: brain conditioning in the *step* fuse order,
    loss vs the oracle, `loss.backward()` -> .grad on the LoRA factors, one optimizer step lowers the loss."""
    from oracle import cs3_dgf as OC
    from oracle import flux_dit as O
    from oracle import train_step as TS
    from loongx_b200.config import FluxConfig
    from src.train.model import OminiModel

    torch.manual_seed(11)
    oc = OC.NeuralConditioner().eval()
    kw = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=4096, pooled_projection_dim=768)
    cfg = FluxConfig(**kw)
    m = OminiModel(cfg, lora_config={"r": 4, "lora_alpha": 4}, device=DEV, model_config={},
                   optimizer_config={"type": "SGD", "params": {"lr": 0.0}}, use_brain_condition=True, fuse_flag=True)
    m.load_state_dict(oc.state_dict(), strict=True)
    oc = oc.to(DEV)
    g = torch.Generator().manual_seed(3)
    B, h, w = 1, 16, 32
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).to(DEV)  # noqa: E731
    batch = dict(image=r(B, 16, h, w).bfloat16(), condition=r(B, 16, h, w).bfloat16(),
                 prompt_embeds=r(B, 512, 4096, scale=0.3).bfloat16(), pooled_prompt_embeds=r(B, 768).bfloat16(),
                 position_delta=[[0, -16]], condition_type=["subject"], t=torch.tensor([0.6]),
                 noise=r(B, 128, 64).bfloat16(), eeg=r(B, 4, 5000), fnirs=r(B, 6, 600), ppg=r(B, 4, 256), motion=r(B, 6, 100))
    loss = m.step(batch)
    assert loss.dim() == 0 and loss.requires_grad
    # oracle with the same (bf16-rounded) DiT weights: pull them back out of the native container
    tr = m._trainer_obj
    ocfg = O.FluxConfig(**kw)
    P = O.init_params(ocfg, seed=0)  # shapes only; overwritten below
    named = m.transformer.weights.named
    # rebuild the oracle's flat dict from the packed panels (qkv panels are stacked in q, k, v(, mlp) order)
    D = ocfg.inner_dim

    def put(name, wt, bias):
        P[name + ".weight"], P[name + ".bias"] = wt.float(), bias.float()

    for key, names in (("x_embedder", ["x_embedder"]), ("context_embedder", ["context_embedder"]),
                       ("time_1", ["time_text_embed.timestep_embedder.linear_1"]),
                       ("time_2", ["time_text_embed.timestep_embedder.linear_2"]),
                       ("guid_1", ["time_text_embed.guidance_embedder.linear_1"]),
                       ("guid_2", ["time_text_embed.guidance_embedder.linear_2"]),
                       ("text_1", ["time_text_embed.text_embedder.linear_1"]),
                       ("text_2", ["time_text_embed.text_embedder.linear_2"]),
                       ("mod_img", ["transformer_blocks.0.norm1.linear"]), ("mod_txt", ["transformer_blocks.0.norm1_context.linear"]),
                       ("mod_single", ["single_transformer_blocks.0.norm.linear"]), ("norm_out", ["norm_out.linear"]),
                       ("proj_out", ["proj_out"]),
                       ("double.0.qkv", ["transformer_blocks.0.attn.to_q", "transformer_blocks.0.attn.to_k", "transformer_blocks.0.attn.to_v"]),
                       ("double.0.qkv_ctx", ["transformer_blocks.0.attn.add_q_proj", "transformer_blocks.0.attn.add_k_proj", "transformer_blocks.0.attn.add_v_proj"]),
                       ("double.0.out", ["transformer_blocks.0.attn.to_out.0"]), ("double.0.out_ctx", ["transformer_blocks.0.attn.to_add_out"]),
                       ("double.0.ff_up", ["transformer_blocks.0.ff.net.0.proj"]), ("double.0.ff_down", ["transformer_blocks.0.ff.net.2"]),
                       ("double.0.ff_ctx_up", ["transformer_blocks.0.ff_context.net.0.proj"]),
                       ("double.0.ff_ctx_down", ["transformer_blocks.0.ff_context.net.2"]),
                       ("single.0.qkv_mlp", ["single_transformer_blocks.0.attn.to_q", "single_transformer_blocks.0.attn.to_k",
                                             "single_transformer_blocks.0.attn.to_v", "single_transformer_blocks.0.proj_mlp"]),
                       ("single.0.proj_out", ["single_transformer_blocks.0.proj_out"])):
        pl, r0 = named[key], 0
        for n in names:
            rows = P[n + ".weight"].shape[0]
            put(n, pl.w[r0:r0 + rows], pl.bias[r0:r0 + rows])
            r0 += rows
    for key in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        P[f"transformer_blocks.0.attn.{key}.weight"] = named[f"double.0.{key}"].float()
    for key in ("norm_q", "norm_k"):
        P[f"single_transformer_blocks.0.attn.{key}.weight"] = named[f"single.0.{key}"].float()
    for n, f in tr.factors.items():
        P[n + ".lora_A.weight"], P[n + ".lora_B.weight"] = f.A.detach().clone(), f.B.detach().clone()
    P = {k: v.to(DEV) for k, v in P.items()}
    ob = {k: (v.float() if isinstance(v, torch.Tensor) and v.dtype == torch.bfloat16 else v) for k, v in batch.items()}
    loss_ref, grads_ref, aux = TS.flow_step_grads(P, ocfg, ob, model_config={}, conditioner=oc, use_brain_condition=True,
                                                  fuse_flag=True)
    l1 = float(loss.detach())
    print(f"\n[OminiModel.step] loss native {l1:.6f} oracle {loss_ref.item():.6f}")
    assert abs(l1 - loss_ref.item()) / loss_ref.item() < 2e-2
    opt = torch.optim.SGD(m.lora_layers, lr=1.0)
    loss.backward()
    got = dict(tr.named_parameters())
    cat_g = torch.cat([got[n].grad.flatten() for n in sorted(got)])
    cat_r = torch.cat([grads_ref[n].flatten() for n in sorted(got)])
    e = _rel(cat_g, cat_r)
    print(f"[OminiModel.step] LoRA .grad relL2 vs oracle autograd {e:.4g}")
    assert e < 6e-2
    # gradient accumulation (accumulate_grad_batches, seed_512.yaml:12): a second backward adds to .grad
    m.step(batch).backward()
    cat_2 = torch.cat([got[n].grad.flatten() for n in sorted(got)])
    assert _rel(cat_2, 2 * cat_g) < 5e-3  # fp32 atomics / TMA reductions are order-dependent: not bitwise reproducible
    for n in got:
        got[n].grad.mul_(0.5)
    for gr in opt.param_groups:
        gr["lr"] = 0.5 / float(cat_g.norm())
    opt.step()
    loss2 = float(m.step(batch).detach())  # step() re-merges the panels whose factors changed
    print(f"[OminiModel.step] loss after optimizer step {loss2:.6f}")
    assert loss2 < l1


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL all-reduce of the LoRA-gradient bucket)")
def test_ddp_gradient_allreduce_two_gpus():
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "scripts", "ddp_train_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(out.stdout[-2000:], out.stderr[-2000:])
    assert out.returncode == 0
    # with the CS3 / DGF conditioning: the bucket carries the encoder gradients too (model.py:656-701, train.py:181-183)
    cmd[cmd.index("29517")] = "29527"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, LX_DDP_BRAIN="1"))
    print(out.stdout[-2000:], out.stderr[-2000:])
    assert out.returncode == 0


def test_micro_batched_step_equals_full_batch():
    """OminiModel.step with micro_batch < B (gradient accumulation inside the step, used when a batch's activations do not
    fit without recompute): same loss and LoRA gradients as the one-shot step."""
    from loongx_b200.config import FluxConfig
    from src.train.model import OminiModel

    kw = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
    g = torch.Generator().manual_seed(5)
    B, h, w = 4, 16, 32
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).bfloat16().to(DEV)  # noqa: E731
    batch = dict(image=r(B, 16, h, w), condition=r(B, 16, h, w), prompt_embeds=r(B, 128, 256, scale=0.5),
                 pooled_prompt_embeds=r(B, 64), position_delta=[[0, -16]], condition_type=["subject"] * B,
                 t=torch.tensor([0.2, 0.4, 0.6, 0.8]), noise=r(B, 128, 64))
    res = {}
    for mb in (None, 2, 1):
        m = OminiModel(FluxConfig(**kw), lora_config={"r": 4, "lora_alpha": 4}, device=DEV, model_config={},
                       use_brain_condition=False, seed=3)
        m.micro_batch = mb
        loss = m.step(batch)
        loss.backward()
        tr = m._trainer_obj
        assert tr.B == (mb or B)
        res[mb] = (float(loss.detach()), torch.cat([p.grad.flatten() for p in tr.parameters()]).clone())
    for mb in (2, 1):
        dl = abs(res[mb][0] - res[None][0]) / res[None][0]
        dg = _rel(res[mb][1], res[None][1])
        print(f"\n[micro-batch {mb}] loss rel diff {dl:.3g}, grad relL2 diff {dg:.3g}")
        assert dl < 1e-3 and dg < 1e-2


def test_optimizer_outlives_trainer_rebuilds_and_lora_scale_changes():
    """configure_optimizers() BEFORE the first step (Lightning's order, model.py:533), then steps at two batch geometries
    and a generate()-style LoRA scale change in between: the optimizer's Parameter objects are the ones every trainer
    differentiates (one stable set per weight set), each step moves the factors and lowers its own loss."""
    from loongx_b200.config import FluxConfig
    from src.train.model import OminiModel

    kw = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=256, pooled_projection_dim=64)
    m = OminiModel(FluxConfig(**kw), lora_config={"r": 4, "lora_alpha": 4}, device=DEV, model_config={},
                   optimizer_config={"type": "SGD", "params": {"lr": 1.0}}, use_brain_condition=False, seed=3)
    opt = m.configure_optimizers()  # no step() yet
    params = [p for grp in opt.param_groups for p in grp["params"]]
    assert len(params) == len(m.lora_layers) and all(a is b for a, b in zip(params, m.lora_layers))
    with torch.no_grad():
        for p in params[1::2]:  # lora_B starts at zero: give the adapters something to differentiate through
            p.normal_(0, 0.02)
    g = torch.Generator().manual_seed(5)
    r = lambda *s, scale=1.0: (torch.randn(*s, generator=g) * scale).bfloat16().to(DEV)  # noqa: E731

    def batch(B, h, w):
        return dict(image=r(B, 16, h, w), condition=r(B, 16, h, w), prompt_embeds=r(B, 128, 256, scale=0.5),
                    pooled_prompt_embeds=r(B, 64), position_delta=[[0, -16]], condition_type=["subject"] * B,
                    t=torch.full((B,), 0.5), noise=r(B, (h // 2) * (w // 2), 64))

    gains = []
    for (B, h, w) in ((2, 16, 32), (1, 16, 16), (2, 16, 32)):
        b = batch(B, h, w)
        m.transformer.set_lora_scale(0.5)  # what generate(joint_attention_kwargs={"scale": .5}) leaves behind
        opt.zero_grad()
        loss = m.step(b)
        assert getattr(m.transformer.weights, "lora_scale", 1.0) == 1.0  # the step runs on scale-1 panels
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)
        assert sum(float(p.grad.abs().sum()) for p in params) > 0
        before = [p.detach().clone() for p in params]
        for p in params:  # per-tensor normalised step (still a descent direction) so that the loss visibly moves
            p.grad.mul_(3e-3 / (float(p.grad.abs().max()) + 1e-30))
        opt.step()
        assert any(not torch.equal(a, p.detach()) for a, p in zip(before, params))
        loss2 = m.step(b)
        print(f"\n[optimizer B={B} {h}x{w}] loss {float(loss.detach()):.6f} -> {float(loss2.detach()):.6f}")
        # (a per-tensor rescaled step is a descent direction only to first order: allow bf16-level noise per case and ask
        # for a net decrease over the three cases)
        assert float(loss2.detach()) < float(loss.detach()) * (1 + 1e-4)
        gains.append(float(loss.detach()) - float(loss2.detach()))
    assert sum(gains) > 0


def test_step_takes_description_strings_when_text_encoders_are_attached():
    """model.py:585-587: `description` is a list of strings (what SeedDataset / collate_step_batch produce); with text
    encoders attached step() encodes them itself, without them it says what is missing."""
    from loongx_b200.config import FluxConfig
    from loongx_b200.text import ClipTextConfig, NativeClipText, NativeT5Encoder, T5Config
    from oracle import text_encoders as T
    from src.flux.pipeline_tools import prepare_text_input
    from src.train.model import OminiModel

    cfg = FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2)
    m = OminiModel(cfg, lora_config={"r": 4, "lora_alpha": 4}, device=DEV, model_config={}, use_brain_condition=False)
    g = torch.Generator().manual_seed(9)
    r = lambda *s: torch.randn(*s, generator=g).bfloat16().to(DEV)  # noqa: E731
    batch = dict(image=r(2, 16, 16, 16), condition=r(2, 16, 16, 16), description=["make it purple", "add a hat"],
                 position_delta=[[0, -8]], condition_type=["subject"] * 2, t=torch.tensor([0.3, 0.7]), noise=r(2, 64, 64))
    with pytest.raises(NotImplementedError, match="text encoders"):
        m.step(batch)
    tk = dict(vocab_size=300, d_model=4096, d_kv=64, num_heads=4, d_ff=512, num_layers=1)
    ck = dict(vocab_size=300, hidden_size=768, intermediate_size=256, num_layers=1, num_heads=12, max_positions=77)
    rnd = lambda P: {k: v.to(torch.bfloat16).float() for k, v in P.items()}  # noqa: E731

    def fake_tokenizer(prompts, padding, max_length, truncation, return_tensors, **kw):
        rows = [[(ord(c) * 7 + i) % 280 + 3 for i, c in enumerate(p[:max_length - 1])] + [299] for p in prompts]
        return {"input_ids": torch.tensor([row + [1] * (max_length - len(row)) for row in rows])}

    m.flux_pipe.attach_text_encoders(clip=NativeClipText(ClipTextConfig(**ck), rnd(T.clip_init(T.ClipCfg(**ck), 8)), "cuda"),
                                     t5=NativeT5Encoder(T5Config(**tk), rnd(T.t5_init(T.T5Cfg(**tk), 7)), "cuda"),
                                     tokenizers=(fake_tokenizer, fake_tokenizer))
    loss_str = float(m.step(batch).detach())
    pe, po, _ = prepare_text_input(m.flux_pipe, batch["description"])
    b2 = {k: v for k, v in batch.items() if k != "description"}
    b2.update(prompt_embeds=pe, pooled_prompt_embeds=po)
    assert loss_str == float(m.step(b2).detach())
