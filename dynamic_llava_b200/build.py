"""Build libdynllava_b200.so (bfloat16) and libdynllava_b200_f16.so (float16, -DDL_HALF) in-tree with nvcc for sm_100a
(cross-compiles without a GPU).  Both come from the same sources and export the same symbols.

    python -m dynamic_llava_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.  cudart is linked statically and libcuda is
never linked (the one driver call, cuTensorMapEncodeTiled, is resolved through cudaGetDriverEntryPoint at run time), so
the library loads — and its symbols can be checked — on a CPU-only machine.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libdynllava_b200.so")
LIB_F16 = os.path.join(CSRC, "libdynllava_b200_f16.so")
SOURCES = [
    "abi.cu",
    "elementwise.cu",
    "select_compact.cu",
    "rope_append.cu",
    "kv_turn.cu",
    "attn_decode.cu",
    "attn_prefill_tc4.cu",
    "skinny_gemm.cu",
    "gemm2.cu",
    "decode_fused.cu",
    "predictor_ops.cu",
    "sample.cu",
    "nocache_select.cu",
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(out: str, deps) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_extension(force: bool = False, verbose: bool = False) -> str:
    """-> path of the bf16 library; the fp16 one is built next to it."""
    _build_one(LIB, "build", [], force, verbose)
    _build_one(LIB_F16, "build_f16", ["-DDL_HALF"], force, verbose)
    return LIB


def _build_one(LIB: str, objdir_name: str, defines, force: bool, verbose: bool) -> str:
    hdrs = [os.path.join(CSRC, "common.cuh"), os.path.join(os.path.dirname(HERE), "include", "dynllava_b200.h")]
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and not _stale(LIB, srcs + hdrs):
        return LIB  # up to date (the GPU box receives the prebuilt .so with the snapshot)
    objdir = os.path.join(CSRC, objdir_name)
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + list(defines) + ["-c", src, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    if force or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


def was_rebuilt(before_mtime) -> bool:
    return (not before_mtime) or max(os.path.getmtime(LIB), os.path.getmtime(LIB_F16)) > before_mtime


if __name__ == "__main__":
    t0 = min(os.path.getmtime(LIB), os.path.getmtime(LIB_F16)) if os.path.exists(LIB) and os.path.exists(LIB_F16) else 0
    path = build_extension(force="--force" in sys.argv, verbose=True)
    print(("built " if was_rebuilt(t0) else "up to date, reusing ") + path + " and " + LIB_F16)
