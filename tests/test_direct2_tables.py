"""The shared-memory rank structure of the DIRECT engine's second form, checked on the CPU: tests/cpp/test_direct2_tables.cpp compiles
the very lookup the kernel runs (csrc/gtb_direct2_tables.h) for the host and compares it with lower_bound over the evaluation points."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_direct2_tables_host(tmp_path):
    exe = str(tmp_path / "t_d2")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "test_direct2_tables.cpp")])
    p = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), p.stdout[-2000:]
