"""ctypes binding of libloongx_b200.so (the C ABI declared in include/loongx_b200.h).

There is deliberately no fallback: if the shared object is missing or does not load, importing this module
raises, so a product path can never silently run on anything but the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os as _os0

# LX_LIB: development aid (A/B builds of the same ABI); the product always loads the in-tree library
_LIB_PATH = Path(_os0.environ["LX_LIB"]) if _os0.environ.get("LX_LIB") else Path(__file__).resolve().parent / "lib" / "libloongx_b200.so"


class LoongXNativeError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not _LIB_PATH.exists():
        raise LoongXNativeError(
            f"{_LIB_PATH} not found: build it with `python -m loongx_b200.build` (needs nvcc); "
            "there is no CPU / PyTorch fallback for the hot path"
        )
    return C.CDLL(str(_LIB_PATH))


lib = _load()

c_void_p, c_int, c_int32, c_int64, c_float = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_float


class TileMeta(C.Structure):
    _fields_ = [("stream", c_int32), ("batch", c_int32), ("seq_row", c_int32), ("reserved", c_int32)]


class GemmSegment(C.Structure):
    _fields_ = [("mode", c_int32), ("col_offset", c_int32), ("out", c_void_p), ("ldo", c_int64)]


class GemmGroup(C.Structure):
    _fields_ = [("W", c_void_p), ("ldw", c_int64), ("bias", c_void_p), ("K", c_int32), ("m_begin", c_int32)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", c_void_p), ("lda", c_int64),
        ("M", c_int32), ("N", c_int32),
        ("n_groups", c_int32), ("n_split", c_int32),
        ("group", GemmGroup * 3),
        ("seg", GemmSegment * 2),
        ("tile_meta", c_void_p),
        ("residual", c_void_p), ("ldr", c_int64),
        ("gate", c_void_p * 3), ("gate_stride", c_int64 * 3),
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p),
        ("heads", c_int32), ("seq_total", c_int32),
        ("rms_q", c_void_p * 3), ("rms_k", c_void_p * 3),
        ("rope", c_void_p),
        ("rms_eps", c_float), ("tile_n", c_int32), ("w_dynamic", c_int32),
        ("col_offset2", c_int32), ("out2", c_void_p), ("ldo2", c_int64),
        ("qkv_pre", c_void_p), ("ld_qkv_pre", c_int64),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p),
        ("out", c_void_p), ("ldo", c_int64),
        ("out_row_base", c_void_p),
        ("col_offset", c_int32),
        ("B", c_int32), ("H", c_int32), ("S", c_int32),
        ("n_cond", c_int32), ("mask_mode", c_int32),
        ("cross_bias", c_float), ("scale", c_float),
        ("lse", c_void_p),
        ("stream_end", c_int32 * 3), ("pad", c_int32 * 3),
        ("q_tiles", c_int32), ("reserved", c_int32),
    ]


class AttnBwdDesc(C.Structure):
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("d_out", c_void_p),
        ("lse", c_void_p), ("delta", c_void_p), ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
        ("B", c_int32), ("H", c_int32), ("S", c_int32), ("n_cond", c_int32), ("mask_mode", c_int32),
        ("cross_bias", c_float), ("scale", c_float), ("reserved", c_int32),
        ("stream_end", c_int32 * 3), ("pad", c_int32 * 3),
    ]


EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_SILU, EPI_GATE_RESIDUAL, EPI_QKV, EPI_BIAS_F32, EPI_BIAS_GELU_DUAL, EPI_MUL_AUX = range(8)

lib.lx_last_error.restype = C.c_char_p
lib.lx_version.restype = c_int
lib.lx_device_info.argtypes = [C.POINTER(c_int32)]
lib.lx_gemm_bf16.argtypes = [C.POINTER(GemmDesc), c_void_p]
lib.lx_attention.argtypes = [C.POINTER(AttnDesc), c_void_p]
lib.lx_attention_bwd.argtypes = [C.POINTER(AttnBwdDesc), c_void_p]
lib.lx_attention_bwd_prep.argtypes = [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                                      c_int32, c_void_p]


import os as _os

if _os.environ.get("LX_PDL"):  # development aid: programmatic dependent launch on (1, default) / off (0)
    lib.lx_debug_set_pdl(int(_os.environ["LX_PDL"]))
if _os.environ.get("LX_ATT_SPLIT"):  # development aid: split-work attention schedule on (1, default) / off (0)
    lib.lx_debug_attention_split(int(_os.environ["LX_ATT_SPLIT"]))
if _os.environ.get("LX_RASTER_MB"):  # development aid: L2 budget of the GEMM raster bands (see gemm.cu::tile_coords)
    lib.lx_debug_gemm_raster_budget_mb(int(_os.environ["LX_RASTER_MB"]))


lib.lx_set_workspace.argtypes = [c_void_p, c_int64, c_void_p]
lib.lx_debug_attention_ctas.argtypes = [c_int]

WORKSPACE_BYTES = 64 << 20  # exchange buffer of the split-work attention / GEMM schedules (lx_set_workspace)
_ws_cache: dict = {}
_ws_key = None


def current_stream() -> int:
    """Raw handle of torch's current CUDA stream.  The first call on a (device, stream) pair allocates the 64 MiB
    exchange workspace of the split-work kernels from the torch allocator and binds it to that stream; the library
    holds one binding at a time, so switching streams re-binds (launches on any other stream run unsplit)."""
    import torch

    global _ws_key
    s = torch.cuda.current_stream()
    key = (s.device.index, s.cuda_stream)
    if key != _ws_key:
        buf = _ws_cache.get(key)
        if buf is None:
            if torch.cuda.is_current_stream_capturing():
                return s.cuda_stream  # no allocation inside a graph capture: unsplit schedule
            buf = torch.zeros(WORKSPACE_BYTES, dtype=torch.uint8, device=s.device)
            _ws_cache[key] = buf
        check(lib.lx_set_workspace(buf.data_ptr(), buf.numel(), s.cuda_stream), "lx_set_workspace")
        _ws_key = key
    return s.cuda_stream


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise LoongXNativeError(f"{what} failed ({rc}): {lib.lx_last_error().decode()}")


def exported_symbols() -> list[str]:
    """Symbols include/loongx_b200.h declares; used by the CPU-side ABI test."""
    import re

    hdr = Path(__file__).resolve().parent.parent / "include" / "loongx_b200.h"
    return sorted(set(re.findall(r"\b(lx_[a-z0-9_]+)\s*\(", hdr.read_text())))
