"""A C host linked against librumdeed_b200.so (-lrumdeed_b200, include/rumdeed_b200.h) -- the link-level stand-in for
the Fortran bridge, which cannot be compiled here: tests/c_boundary/boundary_check.c walks INTEGRATION.md's call
sequence with Fortran-layout arrays and the reference's own unit-test vectors (mod_tests.F90:431-515, :1403-1452)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_boundary", "boundary_check.c")
LIBDIR = os.path.join(ROOT, "rumdeed_b200")


def _build(tmp_path):
    exe = str(tmp_path / "boundary_check")
    # strict C99: the header must be usable from plain C (what ISO_C_BINDING interoperates with)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
           "-L", LIBDIR, "-lrumdeed_b200", "-lm", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_c_host_links_and_refuses_without_gpu(tmp_path):
    """Links on the CPU box; without a device the program must report 'no fallback' (exit 3), never compute."""
    import torch
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3, r.stdout + r.stderr
        assert "no" in r.stdout.lower()


@pytest.mark.gpu
def test_c_host_call_sequence_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "boundary_check ok" in r.stdout
