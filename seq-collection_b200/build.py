"""Builds libfqgpu.so (and the `sc` mirror CLI) in-tree with nvcc for sm_100a.

    python seq-collection_b200/build.py [--force]

The shared library travels to the GPU box with the repo snapshot (it is git-ignored, not
gpurun-ignored).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfqgpu.so")
SC = os.path.join(HERE, "sc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"] + os.environ.get("FQGPU_NVCC_EXTRA", "").split()
LIB_SOURCES = ["fq_scan.cu", "fq_meta.cu", "fqgpu_api.cu", "fq_synth.cu", "fq_shard.cu", "fq_index.cu", "fq_dedup.cu", "fq_bgzf.cu", "fq_gzip.cu"]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps() -> list[str]:
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(HERE, "..", "include", "fqgpu.h"))
    return d


def build_lib(force: bool = False, verbose: bool = False) -> str:
    if force or _stale(LIB, _deps()):
        objs = []
        for src in LIB_SOURCES:
            obj = os.path.join(CSRC, src.replace(".cu", ".o"))
            cmd = [NVCC, *ARCH, *COMMON, "-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            subprocess.run(cmd, check=True)
            objs.append(obj)
        subprocess.run([NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lz", "-ldl"], check=True)
    return LIB


def build_cli(force: bool = False) -> str:
    src = os.path.join(CSRC, "sc_main.cpp")
    if os.path.exists(src) and (force or _stale(SC, [src, LIB])):
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-o", SC, src, "-I", os.path.join(HERE, "..", "include"),
                        "-L", HERE, "-lfqgpu", "-lz", "-Wl,-rpath,$ORIGIN"], check=True)
    return SC


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_lib(force, verbose="-v" in sys.argv))
    print(build_cli(force))
