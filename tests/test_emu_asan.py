"""Bounds check of the kernel's unit code under AddressSanitizer + UBSan (host lane emulation).
compute-sanitizer is closed on the GPU pool, so this is how out-of-range window / input / output
accesses on hostile streams are hunted: same lzgpu_unit.cuh / lzgpu_core.cuh, host pointers."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = "/usr/lib/gcc/x86_64-linux-gnu/13"


@pytest.mark.skipif(not os.path.exists(os.path.join(LIBDIR, "libasan.so")) or not os.path.exists("/usr/bin/g++"),
                    reason="system sanitizer runtime not present")
def test_hostile_inputs_under_asan():
    subprocess.check_call(["make", "-C", os.path.join(HERE, "emu"), "-s", "asan"])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1",
               LD_PRELOAD=f"{LIBDIR}/libasan.so:{LIBDIR}/libubsan.so")
    p = subprocess.run([sys.executable, os.path.join(HERE, "asan_worker.py")], env=env, capture_output=True,
                       text=True, timeout=900)
    assert p.returncode == 0 and "ASAN_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-4000:]
