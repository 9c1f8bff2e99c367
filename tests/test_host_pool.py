"""The engine's host-side helper threads (csrc/host_pool.h), compiled into a small g++ program: no GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_pool(tmp_path):
    exe = str(tmp_path / "host_pool_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "darwin-gpu_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "host_pool_test.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout + out.stderr
