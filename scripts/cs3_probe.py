"""CS3 / DGF conditioning probe (development aid + ncu target): the inference-time neural conditioning of generate()
(generate.py:168-258) on synthetic signals, timed with CUDA events; run under ncu for the per-kernel launch list and the
`--set full` capture of the HBM-bound kernels (profiles/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from loongx_b200 import cs3

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mode = sys.argv[2] if len(sys.argv) > 2 else "all"  # "eeg" = the bench's EEG-only path, "all" = four signals + DGF
dev = "cuda"
torch.manual_seed(0)
eeg_enc, ppg_enc = cs3.EEGEncoder(device=dev), cs3.PPGEncoder(device=dev)
fn_enc, mo_enc = cs3.FNIRSEncoder(device=dev), cs3.MotionEncoder(device=dev)
duan1, duan2 = cs3.DUAN(512, device=dev), cs3.DUAN(1, device=dev)
duan_p, duan_q = cs3.DUAN(512, device=dev), cs3.DUAN(1, device=dev)
fusion1 = torch.nn.Linear(1024, 512).to(dev)
fusion2 = torch.nn.Linear(1536, 768).to(dev)
g = torch.Generator(device=dev).manual_seed(1)
r = lambda *s: torch.randn(*s, generator=g, device=dev)  # noqa: E731
eeg, ppg, fnirs, motion = r(B, 4, 5000), r(B, 4, 256), r(B, 6, 600), r(B, 6, 100)
pe, po = r(B, 512, 4096) * 0.1, r(B, 768)


def run():
    e = eeg_enc(cs3.pad_truncate(eeg, 4096))
    if mode == "eeg":
        return cs3.cast_bf16(e)
    p = ppg_enc(cs3.pad_truncate(ppg, 256))
    cat = torch.empty((B, 1024, 4096), device=dev)
    cat[:, :512].copy_(e)
    duan1(p, e, out=cat[:, 512:])
    tok = cs3.token_axis_linear(fusion1, cat)
    f = fn_enc(cs3.pad_truncate(fnirs, 512))
    m = mo_enc(cs3.pad_truncate(motion, 128))
    cat2 = torch.empty((B, 1536), device=dev)
    cat2[:, :768].copy_(f)
    cat2[:, 768:].copy_(duan2(f.unsqueeze(1), m.unsqueeze(1)).squeeze(1))
    pooled = cs3.gemv(fusion2.weight, fusion2.bias, cat2)
    return cs3.cast_bf16(duan_p(pe, tok)), cs3.cast_bf16(duan_q(po.unsqueeze(1), pooled.unsqueeze(1)))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print(f"B={B} mode={mode}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per conditioning pass")
