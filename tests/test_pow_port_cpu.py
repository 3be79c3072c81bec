"""CPU tests (no GPU) of the bit-exact glibc pow port that the controller uses on the device
(differential-equations_b200/csrc/glibc_pow.h; reference call sites dormandprince/ordinary.rs:154, h_init.rs:124)."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "differential-equations_b200", "csrc")


def test_host_build_of_the_port_matches_libm_bitwise():
    """The very header the kernels compile, built for the host with explicit FMAs only (-ffp-contract=off), returns
    libm's bits on ~2e8 samples: every exponent the path can use, arguments over the whole positive range, near 1,
    subnormal, and the special values."""
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "pow_check")
        subprocess.run(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-I", CSRC,
                        os.path.join(ROOT, "tests", "support", "pow_host_check.cpp"), "-o", exe, "-lpthread"], check=True)
        out = subprocess.run([exe, "500000"], capture_output=True, text=True, check=True)
    total, bad = (int(x) for x in out.stdout.split())
    assert total >= 2.0e8 and bad == 0, out.stderr


def test_committed_tables_are_the_tables_of_this_libm():
    """glibc_pow_tables.h was dumped from the image's libm; the search must find identical data in the libm present."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import io
    import dump_glibc_pow_tables as dump
    buf = io.StringIO()
    dump.emit(dump.find_tables(), buf)
    assert buf.getvalue() == open(os.path.join(CSRC, "glibc_pow_tables.h")).read()
