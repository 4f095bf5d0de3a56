"""The C++ host side above the C ABI (include/lzma_reader.hpp, liblzma_reader.so): the reference's reader API
with its own tests restated in C++ (tests/cpp/reader_test.cpp = reader1_test.go + reader2_test.go + adapters).
CPU tier: it builds, constructors / header errors / helpers behave, and Read fails loudly without a device.
GPU tier: the full program against liblzgpu.so on the B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_build", "reader_test")
ASSETS = os.path.join(ROOT, "tests", "golden", "ref_assets")


@pytest.fixture(scope="module")
def exe():
    # `reader` builds only the C++ host library and this test program (liblzgpu.so must exist: build() made it)
    assert os.path.exists(os.path.join(ROOT, "lzma_b200", "liblzgpu.so")), "run __graft_entry__.build() first"
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "lzma_b200", "csrc"), "-s", "reader"])
    assert os.path.exists(EXE)
    return EXE


def test_cpp_reader_builds_and_fails_loudly_without_device(exe):
    p = subprocess.run([exe, ASSETS, "--no-device"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "PASS" in p.stdout


@pytest.mark.gpu
def test_cpp_reader_reference_tests(exe):
    p = subprocess.run([exe, ASSETS], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    for name in ("TestReader1/bad_file", "TestReader1WithFileVerification", "TestReader2WithFileVerification",
                 "TestSevenZipAdapters", "TestDecodeBatch"):
        assert name in p.stdout, p.stdout
