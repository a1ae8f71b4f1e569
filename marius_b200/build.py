"""marius_b200/build.py -- compile the CUDA library (and the C++/libtorch host adapters) IN-TREE for sm_100a.

    python -m marius_b200.build            # libmarius_b200.so  (+ _host extension when sources exist)

nvcc cross-compiles without a GPU; the built .so files are git-ignored but travel with gpurun snapshots.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

CUDA_SOURCES = ["c_api.cu", "storage_kernels.cu", "radix_sort.cu", "decoder_kernels.cu", "eval_kernels.cu", "sample_kernels.cu", "shard_kernels.cu", "gemm_simt.cu", "gemm_tc_group.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_cuda(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libmarius_b200.so")
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in ("common.cuh", "kernels.h", "decoder_vec.cuh", "gemm_tc_ptx.cuh")] + [os.path.join(HERE, "..", "include", "marius_b200.h")]
    if force or _newer(out, deps):
        # one nvcc process per translation unit (parallel, objects cached under lib/obj), then one link
        objdir = os.path.join(LIBDIR, "obj")
        os.makedirs(objdir, exist_ok=True)
        hdrs = deps[len(srcs):]
        flags = [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("MB_NVCC_EXTRA", "").split()
        objs, procs = [], []
        for s in srcs:
            o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
            objs.append(o)
            if force or _newer(o, [s] + hdrs):
                cmd = [NVCC] + flags + ["-c", s, "-o", o]
                if verbose:
                    print("[marius_b200.build]", " ".join(cmd), flush=True)
                procs.append(subprocess.Popen(cmd, cwd=CSRC))
        if any(p.wait() != 0 for p in procs):
            raise RuntimeError("nvcc failed")
        subprocess.check_call([NVCC, "-shared", "-o", out] + objs, cwd=CSRC)
    return out


def build_host(force: bool = False, verbose: bool = True):
    """C++/libtorch adapters that keep the reference's class surface (Storage, EdgeDecoder, Batch, Model) + pybind module."""
    host_dir = os.path.join(CSRC, "host")
    if not os.path.isdir(host_dir):
        return None
    srcs = sorted(os.path.join(host_dir, f) for f in os.listdir(host_dir) if f.endswith(".cpp"))
    if not srcs:
        return None
    import torch
    from torch.utils import cpp_extension as ce

    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    out = os.path.join(LIBDIR, "_host" + ext)
    hdrs = [os.path.join(host_dir, f) for f in os.listdir(host_dir) if f.endswith(".h")] + [os.path.join(HERE, "..", "include", "marius_b200.h")]
    if force or _newer(out, srcs + hdrs):
        tdir = os.path.dirname(torch.__file__)
        inc = ["-I" + p for p in ce.include_paths()] + ["-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(HERE, "..", "include"),
                                                         "-I/usr/local/cuda/include"]
        import pybind11
        inc.append("-I" + pybind11.get_include())
        objs = []
        objdir = os.path.join(LIBDIR, "obj")
        os.makedirs(objdir, exist_ok=True)
        procs = []
        for s in srcs:
            o = os.path.join(objdir, os.path.basename(s)[:-4] + ".o")
            objs.append(o)
            if force or _newer(o, [s] + hdrs):
                cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-fopenmp", "-w", "-DTORCH_EXTENSION_NAME=_host", "-DTORCH_API_INCLUDE_EXTENSION_H",
                       "-D_GLIBCXX_USE_CXX11_ABI=" + str(int(torch._C._GLIBCXX_USE_CXX11_ABI))] + inc + ["-c", s, "-o", o]
                if verbose:
                    print("[marius_b200.build] CXX", os.path.basename(s), flush=True)
                procs.append(subprocess.Popen(cmd))
        for p in procs:
            if p.wait() != 0:
                raise RuntimeError("host compile failed")
        link = ["g++", "-shared", "-fopenmp", "-o", out] + objs + ["-L" + os.path.join(tdir, "lib"), "-Wl,-rpath," + os.path.join(tdir, "lib"),
                                                                  "-ltorch", "-ltorch_cpu", "-lc10", "-lc10_cuda", "-ltorch_cuda", "-ltorch_python", "-L" + LIBDIR,
                                                                  "-Wl,-rpath,$ORIGIN", "-lmarius_b200"]
        if verbose:
            print("[marius_b200.build] LD", os.path.basename(out), flush=True)
        subprocess.check_call(link)
    return out


def build_all(force: bool = False, verbose: bool = True):
    a = build_cuda(force, verbose)
    b = build_host(force, verbose)
    return a, b


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
