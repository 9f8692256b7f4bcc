"""Runs the C++ host tests (tests/native/test_host.cpp): VoxelMap mirror, indexers, slot arena; with a
GPU also the B200Renderer adapter end to end (frames after brush-like edits == oracle rebuilt from scratch)."""
from __future__ import annotations

import subprocess
from pathlib import Path

import pytest

EXE = Path(__file__).resolve().parent / "native" / "test_host"


def _run(*args):
    if not EXE.exists():
        subprocess.run(["make", "-s", "-C", str(EXE.parent)], check=True)
    r = subprocess.run([str(EXE), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "native tests: ok" in r.stdout, r.stdout + r.stderr
    return r.stdout


def test_host_logic_cpu():
    _run()


@pytest.mark.gpu
def test_adapter_end_to_end_gpu():
    out = _run("--gpu")
    assert "gpu:" in out
