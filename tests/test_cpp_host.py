"""C++ host mirror (include/zstd_b200.hpp: Compressor / Decompressor / ZstdException over the C ABI).

The reference's host side is compiled C# (src/ZstdSharp/Compressor.cs, Decompressor.cs, ThrowHelper.cs); .NET is not in this
image, so the host layer above the C ABI is C++.  tests/cpp/host_mirror_test.cpp restates the behavioural set of
ZstdNetTests.cs against it, with the CPU oracle linked in as the byte-exact checker.  The binary goes to tests/_build/ (git-ignored).
"""
import os
import subprocess

import pytest

from zstdsharp_b200 import _native
from _oracle import build_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_DIR = os.path.join(ROOT, "tests", "_build")
BIN = os.path.join(OUT_DIR, "host_mirror_test")
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")


def build_binary() -> str:
    oracle_lib = build_oracle()
    deps = [SRC, os.path.join(ROOT, "include", "zstd_b200.hpp"), os.path.join(ROOT, "include", "zstd_b200.h")]
    if os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(p) for p in deps):
        return BIN
    os.makedirs(OUT_DIR, exist_ok=True)
    lib_dir, oracle_dir = os.path.dirname(_native.LIB_PATH), os.path.dirname(oracle_lib)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-pthread", SRC, "-o", BIN,
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
           "-L", lib_dir, "-lzstdb200", "-L", oracle_dir, "-lzo",
           "-Wl,-rpath,$ORIGIN/../../zstdsharp_b200/_build", "-Wl,-rpath,$ORIGIN/../../oracle/_build"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return BIN


def run(mode: str) -> str:
    p = subprocess.run([build_binary(), mode], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "checks passed" in p.stdout
    return p.stdout


def test_cpp_host_mirror_compiles_and_host_logic():
    """Header compiles warning-free as C++17, links against the C ABI, host-only behaviour matches the reference's wrapper."""
    run("host")


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu():
    """ZstdNetTests.cs behavioural set through the C++ classes: frames byte-identical to the oracle, round trips, error codes."""
    out = run("gpu")
    assert "gpu:" in out
