"""Build libloongx_b200.so (all CUDA kernels + the C ABI) for sm_100a with nvcc, in-tree.

nvcc cross-compiles without a GPU; the resulting shared object travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
LIB = LIBDIR / "libloongx_b200.so"
STAMP = LIBDIR / ".build_stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=default",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [ROOT.parent / "include" / "loongx_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(exist_ok=True)
    dig = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    objdir = LIBDIR / "obj"
    objdir.mkdir(exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "--shared"]
    for src in sources():
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *flags, "-c", str(src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed on {src.name}\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(f"--- {src.name}\n{out}\n")
    if failed:
        raise RuntimeError("nvcc build failed")
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", str(LIB), *map(str, objs)]  # static cudart: no run-time CUDA toolkit dependency
    subprocess.run(link, check=True)
    STAMP.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
