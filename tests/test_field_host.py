"""CPU check of the field header shared by host and device code (csrc/gl64.cuh): shift-multiplies,
power-of-two roots of unity, add/sub/inverse identities (plonky2 field/src/goldilocks_field.rs)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gl64_host_header():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    with tempfile.TemporaryDirectory() as d:
        exe = os.path.join(d, "field_check")
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "field_check.cpp")])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0, out.stdout + out.stderr
        assert out.stdout.strip().endswith("bad 0")
