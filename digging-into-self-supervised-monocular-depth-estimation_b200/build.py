"""In-tree build of the native code (explicit nvcc / g++ commands, no JIT cache).

  libmd2loss.so   CUDA kernels + the C ABI of include/md2_loss.h      (nvcc, sm_100a only)
  _md2_torch.so   PyTorch C++ extension over that ABI                  (g++, links libmd2loss.so)

Both land next to this file so they travel to the GPU box with the repo snapshot.
nvcc cross-compiles without a GPU.  ``python -m md2_b200.build`` or ``build_all()``.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
ROOT = os.path.dirname(PKG)
LIB = os.path.join(PKG, "libmd2loss.so")
EXT = os.path.join(PKG, "_md2_torch.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd[:3]) + " ...")
    return r.stdout + r.stderr


def build_cuda_library(force=False, verbose=False, ptxas_verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("md2_abi.cu", "md2_l1.cu", "md2_metrics.cu", "md2_pipeline.cu", "md2_jitter.cu", "md2_pad.cu", "md2_pool.cu", "md2_tile.cuh", "md2_platform.h", "md2_nvtx.h",
                                               "md2_host.h")]
    srcs += [os.path.join(ROOT, "include", h) for h in ("md2_loss.h", "md2_ops.h", "md2_metrics.h", "md2_pipeline.h")]
    if not force and _newer(LIB, srcs):
        return LIB
    cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_verbose else []) + srcs[:7] + ["-o", LIB]
    out = _run(cmd, verbose)
    if ptxas_verbose:
        print(out)
    return LIB


def build_torch_extension(force=False, verbose=False):
    src = os.path.join(CSRC, "torch_ext.cpp")
    deps = [src, os.path.join(ROOT, "include", "md2_loss.h"), os.path.join(ROOT, "include", "md2_ops.h")]
    if not force and _newer(EXT, deps) and os.path.exists(LIB):
        return EXT
    import torch
    from torch.utils import cpp_extension
    inc = []
    for p in cpp_extension.include_paths():
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"], "-isystem", os.path.join(CUDA_HOME, "include")]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_md2_torch",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           src, "-o", EXT] + inc + [
        "-L" + tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
        "-L" + os.path.join(CUDA_HOME, "lib64"), "-lcudart",
        "-L" + PKG, "-l:libmd2loss.so",
        "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + tlib, "-Wl,-rpath," + os.path.join(CUDA_HOME, "lib64")]
    _run(cmd, verbose)
    return EXT


def build_all(force=False, verbose=False):
    build_cuda_library(force, verbose)
    build_torch_extension(force, verbose)
    return LIB, EXT


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print("built", LIB, EXT)
