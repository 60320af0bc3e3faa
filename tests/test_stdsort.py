"""mp_stdsort.h re-implements libstdc++'s std::sort (introsort + final insertion sort) so that a kernel can order a read's single-end
seeds exactly as the reference's host code does (singleMerge, DV-DPfunctions.cpp:330-336: an unstable sort whose tie order decides which
seeds survive the 60 % / 200-per-read cut).  This test compiles a checker against the real std::sort."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stdsort_matches_libstdcxx(tmp_path):
    exe = str(tmp_path / "stdsort_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "stdsort_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert out.stdout.startswith("ok")
