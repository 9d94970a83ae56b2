"""Build liblemevit_b200.so in-tree with nvcc for sm_100a (no torch headers, no libcuda link).

    python -m lemevit_b200.build [--force] [-v] [--out=path]

Every translation unit is compiled to its own object (in parallel, only when stale) under
``lemevit_b200/_build/`` and linked into the shared library.  All intermediate and final files are
written under per-process temporary names and renamed atomically, and the whole build runs under an
exclusive file lock, so several ranks that find a stale library at the same time (torchrun, DDP)
cannot hand each other half-written files: the first rank builds, the others wait and find the
library fresh.
"""
from __future__ import annotations

import concurrent.futures as cf
import contextlib
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "liblemevit_b200.so")
SOURCES = ["api.cu", "gemm.cu", "mlp_fused.cu", "tokens.cu", "posembed.cu", "attention_simt.cu", "attention_tc.cu",
           "attention_self.cu", "attention_meta.cu", "dca_fused.cu", "meta_branch.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found; lemevit_b200 needs the CUDA 12.9 toolkit to build its sm_100a kernels")


def _sources():
    return [s for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "lemevit_b200.h"))
    return [h for h in hs if os.path.isfile(h)]


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + _headers()
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


@contextlib.contextmanager
def _locked():
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as f:
        fcntl.flock(f, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(f, fcntl.LOCK_UN)


def _compile_one(nvcc, src, obj, extra, verbose):
    tmp = f"{obj}.{os.getpid()}.tmp"
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", tmp]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        with contextlib.suppress(OSError):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, obj)
    return proc.stderr


def build(force: bool = False, verbose: bool = False, out: str | None = None) -> str:
    lib_path = out or LIB
    if out is None and not force and not _stale():
        return LIB
    nvcc = _nvcc()
    extra = os.environ.get("LMV_NVCC_EXTRA", "").split()   # debug builds, e.g. -DLMV_GEMM_TRACE (tools/gemm_trace.py)
    tag = hashlib.sha1(" ".join(extra).encode()).hexdigest()[:8] if extra else "rel"
    with _locked():
        if out is None and not force and not _stale():      # another process built it while we waited for the lock
            return LIB
        hdr_t = max(os.path.getmtime(h) for h in _headers())
        jobs, objs = [], []
        for s in _sources():
            src = os.path.join(CSRC, s)
            obj = os.path.join(OBJ, f"{os.path.splitext(s)[0]}.{tag}.o")
            objs.append(obj)
            if force or verbose or not os.path.isfile(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
                jobs.append((src, obj))
        logs = []
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as pool:
            futs = [pool.submit(_compile_one, nvcc, src, obj, extra, verbose) for src, obj in jobs]
            for f in futs:
                logs.append(f.result())
        tmp = f"{lib_path}.{os.getpid()}.tmp"
        cmd = [nvcc] + LINK_FLAGS + objs + ["-o", tmp]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
        os.replace(tmp, lib_path)
    if verbose:
        print("\n".join(logs))
    return lib_path


if __name__ == "__main__":
    _out = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=_out[0] if _out else None))
