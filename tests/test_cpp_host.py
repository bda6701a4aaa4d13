"""Builds and runs tests/cpp/host_mirror_test.cpp: the C++ host mirror (include/b2r/registration.hpp) and the PCL
adapter (include/b2r/pcl_adapter.hpp, against tests/cpp/pcl_stub) driven the way the reference's callers drive a
registration object."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", f"-I{ROOT}/include", f"-I{ROOT}/tests/cpp/pcl_stub",
           f"{ROOT}/tests/cpp/host_mirror_test.cpp", "-o", exe, f"-L{ROOT}/mrg_slam_b200", "-lb2r", "-lb2r_synth",
           f"-Wl,-rpath,{ROOT}/mrg_slam_b200", "-L/usr/local/cuda/lib64", "-lcudart", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_cpp_host_mirror_builds_and_refuses_without_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "--expect-gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "gpu path ok" in r.stdout
