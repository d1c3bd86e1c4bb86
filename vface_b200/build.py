"""Build libvface_b200.so (hand-written sm_100a kernels + C-ABI) in-tree with nvcc.

    python -m vface_b200.build [--force] [--verbose]

Cross-compiles without a GPU.  The library lands next to this file so it travels with the repo
snapshot to the GPU box; it is git-ignored (source-only history).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "csrc", "build")
LIB_PATH = os.path.join(HERE, "libvface_b200.so")

SOURCES = ["vf_capi.cu", "vf_ddim.cu", "vf_flow_warp.cu", "vf_fsai.cu", "vf_attn.cu", "vf_attn_f32.cu", "vf_attn_tc.cu", "vf_attn_stream.cu", "vf_attn_pp.cu", "vf_norm.cu", "vf_gemm.cu", "vf_gemm2.cu", "vf_gemm3.cu", "vf_conv_out.cu", "vf_linear.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
] + (os.environ.get("VF_NVCC_EXTRA", "").split())


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _deps_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    paths.append(os.path.join(HERE, "..", "include", "vface_b200.h"))
    paths.append(os.path.abspath(__file__))
    return max(os.path.getmtime(p) for p in paths)


def is_fresh() -> bool:
    return os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _deps_mtime()


def _compile(nvcc: str, src: str, verbose: bool) -> str:
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_fresh():
        return LIB_PATH
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(nvcc, s, verbose), SOURCES))
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-shared", "-o", tmp, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
