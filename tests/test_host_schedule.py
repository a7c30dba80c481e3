"""CPU: the device arithmetic headers compiled for the host (PTX carry ops emulated) against the
C oracle -- checks the 32-bit-limb Montgomery schedule and the XYZZ formulas without a GPU."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_limb_schedule_and_xyzz_against_oracle(tmp_path):
    from oracle import pyoracle
    pyoracle.lib()   # builds liboracle.so if needed
    exe = str(tmp_path / "host_check")
    subprocess.check_call([
        "/usr/bin/g++", "-O2", "-std=c++17", "-x", "c++", os.path.join(ROOT, "tests", "cpu", "host_check.cpp"),
        "-I", os.path.join(ROOT, "simpleworks_b200", "csrc"), "-I", os.path.join(ROOT, "oracle"),
        "-L", os.path.join(ROOT, "oracle"), "-loracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "total mismatches: 0" in out.stdout
