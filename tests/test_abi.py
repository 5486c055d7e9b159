"""CPU-only checks of the drop-in boundary: libfqgpu.so loads, exports every symbol that
include/fqgpu.h declares, its struct layout matches the bindings, and the product fails loudly
(no CPU fallback) when no CUDA device exists."""
import ctypes as C
import os
import re
import subprocess

import pytest

import seq_collection_b200 as fq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fqgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fqgpu_[a-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from importlib import import_module

    build = import_module("seq-collection_b200.build")
    lib_path = build.build_lib()
    assert os.path.exists(lib_path)
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (fqgpu_[a-z_0-9]+)", out))
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in fqgpu.h but not exported: {missing}"
    assert sorted(fq.EXPORTED_SYMBOLS) == declared


def test_struct_layout_and_version():
    lib = fq.load_library()
    assert lib.fqgpu_abi_version() == 1
    assert lib.fqgpu_stats_size() == C.sizeof(fq.Stats)
    assert b"sm_100a" in lib.fqgpu_build_info()
    from oracle import fq_oracle as O

    assert C.sizeof(O.Stats) == C.sizeof(fq.Stats)


def test_sass_is_sm100a():
    """The shipped kernels are native sm_100a code (no PTX-JIT, no other arch)."""
    out = subprocess.run(["cuobjdump", "-lelf", fq.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out)


def test_no_cpu_fallback_without_gpu():
    lib = fq.load_library()
    if lib.fqgpu_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fq.FqGpuError) as ei:
        fq.FqGpu()
    assert ei.value.code == fq.ECUDA
    assert "no CPU fallback" in str(ei.value)
    with pytest.raises(fq.FqGpuError):
        fq.fq_count(os.path.join(ROOT, "tests", "golden", "fastq", "dup.fq"))


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "seq-collection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".nim")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "fq_oracle" not in text and "libfqoracle" not in text, os.path.join(dirpath, f)


def test_host_mirror_formatting():
    assert fq.nim_float_str(0.35) == "0.35"
    assert fq.nim_float_str(1.0) == "1.0"
    assert fq.nim_float_str(14 / 42) == "0.3333333333333333"
    assert fq.nim_float_str(float("nan")) == "nan"
    assert fq.output_header(fq.FQ_COUNT_HEADER, True, False) == "reads\tgc_content\tgc_bases\tn_bases\tbases\tbasename"
    assert fq.output_w_fnames("x", "/a/b/c.fq", True, False) == "x\tc.fq"
