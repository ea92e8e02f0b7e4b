"""loongx_b200 — B200-native (sm_100a) implementation of the LoongX denoising hot path.

The package holds the CUDA kernels + C ABI (csrc/, include/loongx_b200.h) and the thin Python host layer that
mirrors the reference's `src.flux` / `src.train.model` interface.  PyTorch is used for device memory, streams and
torch.distributed only.
"""
__version__ = "0.1.0"
