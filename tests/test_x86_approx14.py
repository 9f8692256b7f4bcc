"""rcp14 / rsqrt14 (the reference's approx_rcp / approx_rsqrt / approx_sqrt, SIMD_AVX512.h:136-138): the product's device/host functions
(voxelrt_b200/csrc/x86_approx14.h) and the oracle's (orc_x86_*) against the INSTRUCTIONS on every one of the 2^32 binary32 inputs.
Needs an AVX-512 host (the build container and the GPU box both are); a few known answers run everywhere."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _has_avx512():
    try:
        return "avx512f" in open("/proc/cpuinfo").read() and "avx512vl" in open("/proc/cpuinfo").read()
    except OSError:
        return False


def test_known_answers():
    from oracle import pyoracle

    lib = pyoracle.load()
    lib.orc_x86_rsqrt14.restype = lib.orc_x86_rcp14.restype = C.c_float
    lib.orc_x86_rsqrt14.argtypes = lib.orc_x86_rcp14.argtypes = [C.c_float]

    def bits(v):
        return int(np.float32(v).view(np.uint32))

    # values read off the instruction (tools/x86_approx14/dump_tables.c): exact at powers of two / four, 14-bit elsewhere
    assert bits(lib.orc_x86_rcp14(1.0)) == 0x3F800000 and bits(lib.orc_x86_rcp14(4.0)) == 0x3E800000
    assert bits(lib.orc_x86_rcp14(np.uint32(0x3F800001).view(np.float32))) == 0x3F7FFE00
    assert bits(lib.orc_x86_rcp14(np.uint32(0x3F800900).view(np.float32))) == 0x3F7FEC80
    assert bits(lib.orc_x86_rsqrt14(1.0)) == 0x3F800000 and bits(lib.orc_x86_rsqrt14(16.0)) == 0x3E800000
    assert bits(lib.orc_x86_rsqrt14(2.0)) == 0x3F350280
    assert bits(lib.orc_x86_rsqrt14(np.uint32(0x3F800001).view(np.float32))) == 0x3F7FFD00
    assert np.isinf(lib.orc_x86_rsqrt14(0.0)) and bits(lib.orc_x86_rsqrt14(-1.0)) == 0xFFC00000
    assert bits(lib.orc_x86_rcp14(float("inf"))) == 0 and bits(lib.orc_x86_rcp14(-0.0)) == 0xFF800000
    rel = [abs(lib.orc_x86_rsqrt14(float(x)) * np.sqrt(x) - 1) for x in np.linspace(0.01, 100, 5000)]
    assert max(rel) < 2.0**-14


@pytest.mark.skipif(not _has_avx512(), reason="needs the rcp14 / rsqrt14 instructions (AVX-512)")
def test_all_2_pow_32_inputs_against_the_instructions(tmp_path):
    exe = tmp_path / "verify_x86_approx14"
    src = ROOT / "tools" / "x86_approx14" / "verify_exhaustive.cpp"
    subprocess.run(["g++", "-O2", "-fopenmp", "-mavx512f", "-mavx512vl", str(src), "-o", str(exe), f"-L{ROOT / 'oracle'}", "-l:liboracle.so",
                    f"-Wl,-rpath,{ROOT / 'oracle'}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "inputs 4294967296 mismatches product rcp14 0 rsqrt14 0 rsqrt14_pos_normal 0 oracle rcp14 0 rsqrt14 0" in r.stdout
