"""ORACLE — test infrastructure only (see oracle/flux_dit.py header for who may import this).

Pure-PyTorch restatement of LoongX's neural-signal conditioning:

  CS3  (cross-scale state-space encoders)   /root/reference/src/train/model.py:16-135 (EEGEncoder), 137-205 (PPGEncoder),
                                             208-274 (FNIRSEncoder), 277-343 (MotionEncoder), 345-373 (FeaturePyramidPooling)
  DGF  (DUAN + fusion linears)              model.py:947-1035 (DUAN), 731-779 (fuse_eeg / fuse_fnirs), 430-454
  length normaliser                          model.py:479-511 (spatial_pyramid_pooling: zero-pad / truncate)
  conditioning glue                          src/flux/generate.py:168-258 (inference), model.py:656-701 (training)

The state-space layer is `from s4torch import S4Model` (model.py:14): a third-party PyPI package that is absent from
/root/reference, un-vendored, un-pinned and not even listed in the requirements files.  Its algorithm is restated here
from the published S4 (DPLR / HiPPO-LegS, "Annotated S4" formulation that s4torch implements) as recorded in
SURVEY.md App. B.  PINNING: every class / method listed above is pinned bit-for-bit to the reference's own source executed
on the CPU (oracle/ref_harness.py -> tests/golden/ref_v1.npz, tests/test_reference_pins_cpu.py: the four encoders,
FeaturePyramidPooling, DUAN(512) / DUAN(1), fuse_eeg / fuse_fnirs, spatial_pyramid_pooling and the generate.py glue
in both fuse modes) — with `S4Model` inside the reference encoders being THIS file's S4Model.  **PARITY UNPINNED** for
the S4 layer itself: nothing in the reference pins s4torch's numbers; the only available cross-check
(tests/test_oracle_cpu.py) is that the Cauchy/iFFT convolution kernel equals the bilinear-discretised recurrence.

Documented deviations from the literal reference (all required for the path to run at all):
  D1  encoders take [B, C, L] (generate.py passes .flatten(1), which crashes in EEGEncoder.forward)
  D3/D4  everything here runs in fp32 (the reference only works with dtype float32, train/config/seed_512.yaml:2)
  D2  both fuse orders exist: `conditioning(..., mode="generate")` = generate.py:240-258, mode="step" = model.py:680-698
  D5  EEG-only: literal behaviour leaves the embeddings untouched; `eeg_only_replace=True` opts in to replacing prompt_embeds
  D6  batched [B, C, L] signals are accepted
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------------------------
# S4
# ------------------------------------------------------------------------------------------------------------------
def make_hippo(n: int) -> torch.Tensor:
    """HiPPO-LegS matrix (positive convention), 1-indexed: A[i,k] = sqrt(2i+1)sqrt(2k+1) (i>k), i+1 (i==k), 0 (i<k)."""
    idx = torch.arange(1, n + 1, dtype=torch.float64)
    pre = torch.sqrt(2 * idx + 1)
    a = torch.tril(pre[:, None] * pre[None, :], diagonal=-1) + torch.diag(idx + 1)
    return a


def make_nplr(n: int):
    """A = -HiPPO; p = 0.5 sqrt(2i+1), q = 2p; diagonalise S = A + p q^T = V diag(lambda) V^H; return (lambda, V^H p, V^H q)."""
    nhippo = -make_hippo(n)
    p = 0.5 * torch.sqrt(2 * torch.arange(1, n + 1, dtype=torch.float64) + 1.0)
    q = 2 * p
    s = nhippo + p[:, None] * q[None, :]
    lam, v = torch.linalg.eig(s)
    vc = v.conj().T
    return lam, vc @ p.to(torch.complex128), vc @ q.to(torch.complex128)


def s4_kernel(lam, p, q, B, Ct, log_step, L: int) -> torch.Tensor:
    """DPLR generating function at the L roots of unity + inverse FFT -> real conv kernel [d_model, L].
    lam, p, q: [n] complex; B, Ct: [d, n] complex; log_step: [d]."""
    cdtype = lam.dtype
    rdtype = torch.float64 if cdtype == torch.complex128 else torch.float32
    step = torch.exp(log_step.to(rdtype))[:, None]  # [d,1]
    omega = torch.exp(-2j * math.pi * torch.arange(L, device=lam.device, dtype=rdtype) / L).to(cdtype)  # [L]
    g = (2.0 / step) * ((1.0 - omega) / (1.0 + omega))[None, :]  # [d, L]
    c = (2.0 / (1.0 + omega))[None, :]
    a0, a1 = Ct.conj(), q.conj()[None, :].expand_as(Ct)
    b0, b1 = B, p[None, :].expand_as(B)
    denom = g[:, :, None] - lam[None, None, :]  # [d, L, n]

    def cauchy(a, b):
        return ((a * b)[:, None, :] / denom).sum(-1)

    k00, k01, k10, k11 = cauchy(a0, b0), cauchy(a0, b1), cauchy(a1, b0), cauchy(a1, b1)
    at_roots = c * (k00 - k01 * (1.0 / (1.0 + k11)) * k10)
    return torch.fft.ifft(at_roots, n=L, dim=-1).real


class S4Layer(nn.Module):
    def __init__(self, d_model: int, n: int, l_max: int):
        super().__init__()
        self.d_model, self.n, self.l_max = d_model, n, l_max
        lam, p, q = make_nplr(n)
        self.register_buffer("lambda_", lam.to(torch.complex64))
        self.register_buffer("p", p.to(torch.complex64))
        self.register_buffer("q", q.to(torch.complex64))
        self.B = nn.Parameter(nn.init.xavier_normal_(torch.empty(d_model, n)).to(torch.complex64))
        self.Ct = nn.Parameter(nn.init.xavier_normal_(torch.empty(d_model, n)).to(torch.complex64))
        self.D = nn.Parameter(torch.ones(1, 1, d_model))
        self.log_step = nn.Parameter(torch.rand(d_model) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))

    def kernel(self, L: Optional[int] = None) -> torch.Tensor:
        return s4_kernel(self.lambda_, self.p, self.q, self.B, self.Ct, self.log_step, L or self.l_max)

    def forward(self, u: torch.Tensor) -> torch.Tensor:  # [B, L, d]
        L = u.shape[1]
        K = self.kernel(L)  # [d, L]
        ud = torch.fft.rfft(u.float(), n=2 * L, dim=1)
        kd = torch.fft.rfft(K.T.float(), n=2 * L, dim=0)[None]
        y = torch.fft.irfft(ud * kd, n=2 * L, dim=1)[:, :L]
        return y + self.D * u


class S4Block(nn.Module):
    """post-LayerNorm residual block: LN(u + Linear(GELU(S4Layer(u)))) (dropout p = 0)."""

    def __init__(self, d_model: int, n: int, l_max: int):
        super().__init__()
        self.s4 = S4Layer(d_model, n, l_max)
        self.linear = nn.Linear(d_model, d_model)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, u):
        return self.norm(u + self.linear(F.gelu(self.s4(u))))


class S4Model(nn.Module):
    def __init__(self, d_input: int, d_model: int, d_output: int, n_blocks: int, n: int, l_max: int):
        super().__init__()
        self.encoder = nn.Linear(d_input, d_model)
        self.blocks = nn.ModuleList([S4Block(d_model, n, l_max) for _ in range(n_blocks)])
        self.decoder = nn.Linear(d_model, d_output)

    def forward(self, u):  # [B, L, d_input] -> [B, L, d_output]
        y = self.encoder(u)
        for blk in self.blocks:
            y = blk(y)
        return self.decoder(y)


# ------------------------------------------------------------------------------------------------------------------
# CS3 encoders (model.py:16-373).  Dropout(0.3) is identity in eval and is omitted; Sequential indices are kept so the
# state-dict keys equal the reference's (projection.1 / .2 / .5 / .6 / .10).
# ------------------------------------------------------------------------------------------------------------------
class FeaturePyramidPooling(nn.Module):
    def __init__(self, output_sizes):
        super().__init__()
        self.output_sizes = list(output_sizes)

    def forward(self, x):  # [B, C, L]
        return torch.cat([F.adaptive_avg_pool1d(x, s) for s in self.output_sizes], dim=-1)


def _mlp_tokens(d_in, d_hid):
    return nn.Sequential(
        nn.Flatten(start_dim=1), nn.Linear(d_in, d_hid), nn.LayerNorm(d_hid), nn.ReLU(), nn.Identity(),
        nn.Linear(d_hid, 4096), nn.LayerNorm(4096), nn.ReLU(), nn.Identity(), nn.Unflatten(1, (512, 8)), nn.Linear(8, 4096))


def _mlp_pooled(d_in, d_hid):
    return nn.Sequential(
        nn.Flatten(start_dim=1), nn.Linear(d_in, d_hid), nn.LayerNorm(d_hid), nn.ReLU(), nn.Identity(),
        nn.Linear(d_hid, 768), nn.LayerNorm(768), nn.ReLU(), nn.Identity())


class EEGEncoder(nn.Module):  # model.py:16-135
    def __init__(self):
        super().__init__()
        self.eeg_fixed_length = 4096
        self.s41 = S4Model(4, 64, 64, 2, 64, 4096)
        self.s42 = S4Model(4, 4, 4, 2, 4, 4096)
        self.fpp = FeaturePyramidPooling([128, 256, 512, 1024, 2048])
        self.projection = _mlp_tokens(4 * 4096, 2048)

    def forward(self, x):  # [B, 4, 4096] -> [B, 512, 4096]
        z1 = self.s41(x.permute(0, 2, 1))  # [B, L, 64]
        z1 = F.adaptive_avg_pool1d(z1.permute(0, 2, 1), 4).permute(0, 2, 1)  # [B, 4, 64]
        z2 = self.s42(x.permute(0, 2, 1))
        z2 = F.adaptive_avg_pool1d(z2.permute(0, 2, 1), 64)  # [B, 4, 64]
        return self.projection(torch.cat([z1, self.fpp(x), z2], dim=-1))


class _SmallEncoder(nn.Module):
    def __init__(self, ch, length, pool, fpp_sizes, d_hid, tokens: bool):
        super().__init__()
        self.s4 = S4Model(ch, ch, ch, 2, ch, length)
        self.pool = pool
        self.fpp = FeaturePyramidPooling(fpp_sizes)
        d_in = ch * pool + ch * sum(fpp_sizes)
        self.projection = _mlp_tokens(d_in, d_hid) if tokens else _mlp_pooled(d_in, d_hid)

    def forward(self, x):
        z = self.s4(x.permute(0, 2, 1)).permute(0, 2, 1)
        z = F.adaptive_avg_pool1d(z, self.pool)
        return self.projection(torch.cat([z.flatten(1), self.fpp(x).flatten(1)], dim=1))


class PPGEncoder(_SmallEncoder):  # model.py:137-205  [B,4,256] -> [B,512,4096]
    def __init__(self):
        super().__init__(4, 256, 16, [64, 128, 256], 1024, True)


class FNIRSEncoder(_SmallEncoder):  # model.py:208-274  [B,6,512] -> [B,768]
    def __init__(self):
        super().__init__(6, 512, 32, [128, 256, 448], 1024, False)


class MotionEncoder(_SmallEncoder):  # model.py:277-343  [B,6,128] -> [B,768]
    def __init__(self):
        super().__init__(6, 128, 6, [32, 64, 124], 512, False)


def spatial_pyramid_pooling(x: torch.Tensor, output_size: int, adaptive: bool = False) -> torch.Tensor:
    """model.py:479-511: despite the name, zero-pad or truncate the last dim (adaptive=True: adaptive_avg_pool1d)."""
    B, Cc, length = x.shape
    if length == output_size:
        return x
    if adaptive:
        return F.adaptive_avg_pool1d(x, output_size)
    if length < output_size:
        return torch.cat([x, torch.zeros(B, Cc, output_size - length, device=x.device, dtype=x.dtype)], dim=2)
    return x[:, :, :output_size]


# ------------------------------------------------------------------------------------------------------------------
# DGF (model.py:947-1035)
# ------------------------------------------------------------------------------------------------------------------
class DUAN(nn.Module):
    def __init__(self, channels: int, hidden_dim: int = 128, keep_ratio: float = 0.7, eps: float = 1e-3):
        super().__init__()
        self.channels, self.hidden_dim, self.keep_ratio, self.eps = channels, hidden_dim, keep_ratio, eps
        self.gate = nn.Sequential(nn.Conv1d(channels, hidden_dim, 1), nn.ReLU(), nn.Conv1d(hidden_dim, channels, 1),
                                  nn.Sigmoid())
        self.mlp = nn.Sequential(nn.Conv1d(channels, hidden_dim, 1), nn.ReLU(), nn.Conv1d(hidden_dim, channels * 2, 1))

    def forward(self, x16, c16, keep_ratio=None, return_aux: bool = False):
        x, c = x16.float(), c16.float()
        assert x.shape == c.shape
        B, Cc, Ln = x.shape
        keep_ratio = self.keep_ratio if keep_ratio is None else keep_ratio
        mu_c = x.mean(dim=2, keepdim=True)
        sigma_c = torch.sqrt(x.var(dim=2, unbiased=False, keepdim=True) + self.eps)
        mu_l = x.mean(dim=(1, 2), keepdim=True).expand(B, Cc, 1)
        sigma_l = torch.sqrt(x.var(dim=(1, 2), unbiased=False, keepdim=True).expand(B, Cc, 1) + self.eps)
        g_mix = self.gate(c).mean(dim=2, keepdim=True)
        mu = g_mix * mu_c + (1 - g_mix) * mu_l
        sigma = g_mix * sigma_c + (1 - g_mix) * sigma_l
        x_hat = (x - mu) / sigma
        gamma, beta = self.mlp(c.mean(dim=2, keepdim=True)).chunk(2, dim=1)
        y = (1 + gamma) * x_hat + beta
        imp = y.abs().mean(dim=2)
        k = max(1, int(Cc * keep_ratio))
        topk = torch.topk(imp, k, dim=1).indices
        mask = torch.zeros_like(imp).scatter_(1, topk, 1.0)
        y = y * mask.unsqueeze(2)
        out = y.to(x16.dtype)
        return (out, imp, mask) if return_aux else out


class NeuralConditioner(nn.Module):
    """The CS3/DGF sub-modules of OminiModel (model.py:417-462) with the reference attribute names."""

    def __init__(self):
        super().__init__()
        self.eeg_fixed_length, self.fnirs_fixed_length, self.ppg_fixed_length, self.motion_fixed_length = 4096, 512, 256, 128
        self.fusion1 = nn.Sequential(nn.Linear(512 * 2, 512))
        self.fusion2 = nn.Sequential(nn.Linear(768 + 768, 768))
        self.duan_norm1 = DUAN(512)
        self.duan_norm2 = DUAN(1)
        self.fusion3 = nn.Sequential(nn.Linear(512 * 2, 512))
        self.fusion4 = nn.Sequential(nn.Linear(768 * 2, 768))
        self.duan_norm_prompt = DUAN(512)
        self.duan_norm_pooled = DUAN(1)
        self.eeg_projection = EEGEncoder()
        self.ppg_projection = PPGEncoder()
        self.fnirs_projection = FNIRSEncoder()
        self.motion_projection = MotionEncoder()

    def fuse_eeg(self, eeg_features, ppg_features):  # model.py:731-755
        fused = self.duan_norm1(ppg_features, eeg_features)
        fused = torch.cat([eeg_features, fused], dim=1).transpose(1, 2).contiguous()
        return self.fusion1(fused).transpose(1, 2).contiguous()

    def fuse_fnirs(self, fnirs_features, motion_features):  # model.py:757-779
        f, m = fnirs_features.unsqueeze(1), motion_features.unsqueeze(1)
        fused = self.duan_norm2(f, m)
        return self.fusion2(torch.cat([f, fused], dim=-1)).squeeze(1)

    def brain_embeddings(self, eeg=None, fnirs=None, ppg=None, motion=None):
        """generate.py:168-237 / model.py:628-678: raw [B,C,L] signals -> (prompt_embeds_brain, pooled_brain)."""
        pe_b = po_b = None
        if eeg is not None:
            e = self.eeg_projection(spatial_pyramid_pooling(eeg.float(), self.eeg_fixed_length))
            if ppg is not None:
                p = self.ppg_projection(spatial_pyramid_pooling(ppg.float(), self.ppg_fixed_length))
                pe_b = self.fuse_eeg(e, p)
            else:
                pe_b = e
        if fnirs is not None:
            f = self.fnirs_projection(spatial_pyramid_pooling(fnirs.float(), self.fnirs_fixed_length))
            if motion is not None:
                m = self.motion_projection(spatial_pyramid_pooling(motion.float(), self.motion_fixed_length))
                po_b = self.fuse_fnirs(f, m)
            else:
                po_b = f
        return pe_b, po_b

    def conditioning(self, prompt_embeds, pooled, eeg=None, fnirs=None, ppg=None, motion=None, fuse_flag=True,
                     mode="generate", eeg_only_replace=False):
        """-> (prompt_embeds, pooled) after the neural conditioning, in the input dtype."""
        pe_b, po_b = self.brain_embeddings(eeg, fnirs, ppg, motion)
        dt = prompt_embeds.dtype
        if pe_b is not None and po_b is not None:
            if fuse_flag and mode == "generate":  # generate.py:240-255
                pe = self.duan_norm_prompt(prompt_embeds.float(), pe_b)
                po = self.duan_norm_pooled(pooled.float().unsqueeze(1), po_b.unsqueeze(1)).squeeze(1)
                return pe.to(dt), po.to(dt)
            if fuse_flag and mode == "step":  # model.py:680-698
                pe32, po32 = prompt_embeds.float(), pooled.float()
                fused = self.duan_norm_prompt(pe_b, pe32)
                cat = torch.cat([pe32, fused], dim=1).transpose(1, 2).contiguous()
                pe = pe32 + self.fusion3(cat).transpose(1, 2).contiguous()
                fp = self.duan_norm_pooled(po_b.unsqueeze(1), po32.unsqueeze(1))
                po = po32 + self.fusion4(torch.cat([po32, fp.squeeze(1)], dim=-1))
                return pe.to(dt), po.to(dt)
            return pe_b.to(dt), po_b.to(dt)  # generate.py:256-258 / model.py:699-701
        if eeg_only_replace and pe_b is not None:  # D5 opt-in
            return pe_b.to(dt), pooled
        return prompt_embeds, pooled
