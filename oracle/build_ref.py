#!/usr/bin/env python
"""Build the UNMODIFIED reference CUDA extensions for sm_100a into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path.

The reference's own build (`kernels/setup.py`) refuses compute capability 10.0
(kernels/setup.py:16,76-83), so this recipe drives nvcc/g++ directly on the
sources where they lie under /root/reference (never copied into this repo) with
the reference's own flags (kernels/setup.py:19-36) and only the arch changed to
`-gencode arch=compute_100a,code=sm_100a`.  Outputs go to oracle/_ref/ (git-ignored,
but shipped to the GPU box by gpurun).  Each module is renamed
`ref_<name>` through -DTORCH_EXTENSION_NAME so it can be imported next to the
product's `qserve_backend.<name>`.

On the GPU box the built modules serve as (1) a live parity reference for the
`-m gpu` tests and (2) the "legacy mma.sync on B200" timing row of bench.py.

Usage: python oracle/build_ref.py [name ...]   (default: all)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

REF = "/root/reference/kernels/csrc"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")

# name -> sources (relative to kernels/csrc), mirrors kernels/setup.py:158-245
EXTS = {
    "qgemm_w4a8_per_chn": ["qgemm/w4a8_per_chn/pybind.cpp", "qgemm/w4a8_per_chn/gemm_cuda.cu"],
    "qgemm_w4a8_per_group": ["qgemm/w4a8_per_group/pybind.cpp", "qgemm/w4a8_per_group/gemm_cuda.cu"],
    "qgemm_w8a8": ["qgemm/w8a8/pybind.cpp", "qgemm/w8a8/w8a8_gemm_cuda.cu"],
    "fused_kernels": ["fused.cpp", "fused_kernels.cu"],
    "layernorm_ops": ["layernorm.cpp", "layernorm_kernels.cu"],
    "activation_ops": ["activation.cpp", "activation_kernels.cu"],
    "fused_attention": [
        "fused_attention/fused_attention.cpp",
        "fused_attention/decoderMaskedMultiheadAttention.cu",
        "fused_attention/update_kv_cache.cu",
        "fused_attention/input_metadata_helper.cu",
    ],
}


def _flags():
    import torch
    from torch.utils import cpp_extension as ce

    abi = 1 if torch._C._GLIBCXX_USE_CXX11_ABI else 0
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}", "-I/usr/local/cuda/include"]
    common = ["-std=c++17", "-DENABLE_BF16", f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    cxx = ["-g0", "-O3", "-fopenmp", "-fPIC"] + common
    nvcc = [
        "-O2",
        "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
        "-U__CUDA_NO_BFLOAT16_OPERATORS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
        "-U__CUDA_NO_BFLOAT162_OPERATORS__", "-U__CUDA_NO_BFLOAT162_CONVERSIONS__",
        "--expt-relaxed-constexpr", "--expt-extended-lambda", "--use_fast_math",
        "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-w",
    ] + common
    libdir = ce.library_paths()[0]
    link = [f"-L{libdir}", "-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
            "-L/usr/local/cuda/lib64", "-lcudart", "-lgomp", f"-Wl,-rpath,{libdir}"]
    return inc, cxx, nvcc, link


def _compile(args):
    name, src, inc, cxx, nvcc = args
    path = os.path.join(REF, src)
    obj = os.path.join(OBJ, name + "__" + src.replace("/", "_") + ".o")
    if os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(path):
        return obj
    define = [f"-DTORCH_EXTENSION_NAME=ref_{name}"]
    if src.endswith(".cu"):
        cmd = ["nvcc", "-c", path, "-o", obj, "--threads", "2"] + nvcc + inc + define
    else:
        cmd = ["g++", "-c", path, "-o", obj] + cxx + inc + define
    print("[build_ref]", " ".join(cmd[:6]), "...", flush=True)
    subprocess.check_call(cmd)
    return obj


def main(names):
    os.makedirs(OBJ, exist_ok=True)
    inc, cxx, nvcc, link = _flags()
    jobs = [(n, s, inc, cxx, nvcc) for n in names for s in EXTS[n]]
    with ThreadPoolExecutor(max_workers=int(os.environ.get("REF_BUILD_JOBS", "3"))) as ex:
        objs = list(ex.map(_compile, jobs))
    by_name = {}
    for (n, *_), o in zip(jobs, objs):
        by_name.setdefault(n, []).append(o)
    for n, os_ in by_name.items():
        so = os.path.join(OUT, f"ref_{n}.so")
        subprocess.check_call(["g++", "-shared", "-o", so] + os_ + link)
        print("[build_ref] built", so, flush=True)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        print("[build_ref] /root/reference not present; nothing to do")
        sys.exit(0)
    main(sys.argv[1:] or list(EXTS))
