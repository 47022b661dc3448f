"""The device math headers of the CUDA path (piclas_b200/csrc/math.cuh, fastmath.cuh) compiled for the host
(tests/device_math_host.cpp): restructured vs reference-order Lagrange basis, field evaluation and push on 1e5 random inputs.
CPU only; it checks the arithmetic the kernels are built from, the -m gpu parity tests check the kernels."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_restructured_device_arithmetic_agrees_with_reference_order(tmp_path):
    exe = str(tmp_path / "device_math_host")
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-I", cuda_inc, "-I", os.path.join(ROOT, "piclas_b200", "csrc"),
                    "-o", exe, os.path.join(ROOT, "tests", "device_math_host.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("lagrange")
