"""Time the native VAE at BASELINE's image size (512 x 512; LX_VAE_PX overrides) with CUDA events: decode and encode,
per-class launch totals from the library's profiler.  Usage: python scripts/vae_probe.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loongx_b200 import _lib as L  # noqa: E402
from loongx_b200.vae import NativeVae, VaeConfig, VaeWeights, synthetic_params  # noqa: E402


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    px = int(os.environ.get("LX_VAE_PX", 512))
    cfg = VaeConfig()
    v = NativeVae(VaeWeights(cfg, synthetic_params(cfg), "cuda"))
    g = torch.Generator(device="cuda").manual_seed(0)
    z = torch.randn(B, 16, px // 8, px // 8, generator=g, device="cuda")
    img = torch.rand(B, 3, px, px, generator=g, device="cuda") * 2 - 1
    a, b = v.decode(z).sample, v.decode(z).sample
    print(f"[vae determinism] LX_PDL={os.environ.get('LX_PDL', '1')}: two decodes bit-identical: {torch.equal(a, b)}, "
          f"relL2 {((a - b).norm() / b.norm()).item():.3g}")
    for name, fn in (("decode", lambda: v.decode(z)), ("encode", lambda: v.encode(img).latent_dist.mode())):
        ms = timed(fn)
        v.launches = 0
        fn()
        import ctypes as C

        L.lib.lx_profile_begin()
        fn()
        t, n, w = (C.c_double * 4)(), (C.c_int64 * 4)(), (C.c_double * 4)()
        L.lib.lx_profile_end(t, n, w)
        print(f"[vae {name}] B={B} {px}x{px}: {ms:.2f} ms/call, {v.launches} launches; "
              f"GEMM {t[0]:.2f} ms ({w[0] / max(t[0], 1e-9) / 1e9:.0f} TFLOP/s over {n[0]} launches), "
              f"row kernels {t[2]:.2f} ms ({w[2] / max(t[2], 1e-9) / 1e6:.0f} GB/s over {n[2]} launches)")


if __name__ == "__main__":
    main()
