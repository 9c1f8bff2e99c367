"""Host-side units of the engine compiled into small programs: the helper thread pool (csrc/host_pool.h) and the inter-task
kernel band geometry (csrc/gact_kernels_it.cuh).  No GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_pool(tmp_path):
    exe = str(tmp_path / "host_pool_test")
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "darwin-gpu_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "host_pool_test.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout + out.stderr


def test_inter_task_band_geometry(tmp_path):
    """it_geometry() of csrc/gact_kernels_it.cuh on the host (nvcc compiles the program, nothing runs on a GPU)."""
    exe = str(tmp_path / "it_geometry_test")
    subprocess.run(["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                    "-I", os.path.join(ROOT, "darwin-gpu_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "it_geometry_test.cu"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr
