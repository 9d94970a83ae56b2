"""Build liblemevit_b200.so in-tree with nvcc for sm_100a (no torch headers, no libcuda link).

    python -m lemevit_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblemevit_b200.so")
SOURCES = ["api.cu", "gemm.cu", "mlp_fused.cu", "tokens.cu", "posembed.cu", "attention_simt.cu", "attention_tc.cu", "attention_self.cu", "attention_meta.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "--use_fast_math=false", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; lemevit_b200 needs the CUDA 12.9 toolkit to build its sm_100a kernels")


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), "include", "lemevit_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    if out is not None:
        force = True
    lib_path = out or LIB
    if not force and not _stale():
        return LIB
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += os.environ.get("LMV_NVCC_EXTRA", "").split()   # debug builds, e.g. -DLMV_GEMM_TRACE (tools/gemm_trace.py)
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", lib_path + ".tmp"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(lib_path + ".tmp", lib_path)
    if verbose:
        print(proc.stderr)
    return lib_path


if __name__ == "__main__":
    _out = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out[0] if _out else None))
