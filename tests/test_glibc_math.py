"""The exp / log transcription post_sw runs on the device (shrimp_b200/csrc/glibc_math.cuh), compiled for the host,
against this machine's libm: every bit must agree, or post_sw's base calls would differ from the reference's in the
columns where equally likely alternatives tie.  tools/check_glibc_math.c is the comparison."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_transcription_equals_host_libm(tmp_path):
    exe = str(tmp_path / "chk")
    subprocess.run(["g++", "-O2", "-x", "c++", os.path.join(ROOT, "tools", "check_glibc_math.c"), "-o", exe, "-lm"],
                   check=True, capture_output=True)
    r = subprocess.run([exe, "5000000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "0 mismatches" in r.stdout


def test_tables_are_the_installed_libms(tmp_path):
    """the committed tables equal what tools/gen_glibc_tables.py extracts from the libm of this image"""
    out = str(tmp_path / "t.inc")
    r = subprocess.run(["python", os.path.join(ROOT, "tools", "gen_glibc_tables.py"), out], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("another libm build: " + (r.stderr.strip().splitlines() or ["?"])[-1])
    assert open(out).read() == open(os.path.join(ROOT, "shrimp_b200", "csrc", "glibc_tables.inc")).read()
