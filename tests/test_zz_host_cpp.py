"""The compiled host above the C ABI (piclas_b200/host): the reference's host is compiled Fortran, the image has no Fortran
compiler, so the call sequence of TimeStepPoissonByBorisLeapfrog is mirrored in C++ (particle_step.hpp) and driven by
run_case.cpp from a case file holding what the Fortran host owns at run time (tables, parameters, particles, field).

CPU: the driver builds warning-free, its field tables follow the header, its file format round-trips, and without a CUDA
device it aborts with the library's message (the product path has no CPU fallback).
GPU: the C++ driver and the Python host (ctypes) run the same case through the same library and must agree bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import cases
from piclas_b200 import build, casefile, hostmesh as hm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def _case(tmp_path, nsteps=3):
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (4, 3, 5), 2)
    prm = cases.electron_params(arithmetic=1)
    dt = 1e-8
    PS, spec = cases.uniform_plasma(mesh, 6000, seed=4, vth_cells=0.4, dt=dt)
    elem = hm.cartesian_locate(mesh, PS[:, :3])
    E = cases.smooth_field(mesh, amp=1e-4)
    ids = np.arange(len(spec), dtype=np.int64)
    path = str(tmp_path / "case.bin")
    casefile.write_case(path, mesh, prm, PS, spec, elem, E, dt, nsteps, ids=ids)
    return path, mesh, prm, PS, spec, elem, E, dt, ids


def test_field_tables_follow_the_header():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_fields", os.path.join(ROOT, "piclas_b200", "host", "gen_fields.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    assert open(os.path.join(ROOT, "piclas_b200", "host", "pgpu_fields.inc")).read() == gen.generate(), \
        "run python piclas_b200/host/gen_fields.py"


def test_driver_builds_and_round_trips_the_case_format(tmp_path):
    build.build_cuda()
    exe = build.build_host(force=True)                     # -Wall -Wextra -Werror
    path = _case(tmp_path)[0]
    copy = str(tmp_path / "copy.bin")
    r = subprocess.run([exe, "--echo", path, copy], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = casefile.read_arrays(path), casefile.read_arrays(copy)
    assert list(a) == list(b) and all(np.array_equal(a[k], b[k]) and a[k].dtype == b[k].dtype for k in a)
    assert open(path, "rb").read() == open(copy, "rb").read()
    # every scalar / array field of both structs is in the file; absent pointers are NULL in the C++ structs
    names = set(a)
    for f in ("mesh.nGlobalElems", "mesh.ElemInfo", "mesh.NodeVolume", "mesh.xyzminglob", "params.TrackingMethod",
              "params.externalField", "params.ChargeIC", "params.maxParticleNumber", "part.PartState", "field.E", "run.dt"):
        assert f in names, f
    assert "mesh.SFElemr2" not in names


@pytest.mark.skipif(_has_gpu(), reason="checks the behaviour without a CUDA device")
def test_driver_aborts_with_the_librarys_message_without_a_device(tmp_path):
    build.build_cuda()
    exe = build.build_host()
    path = _case(tmp_path)[0]
    r = subprocess.run([exe, path, str(tmp_path / "res.bin")], capture_output=True, text=True)
    assert r.returncode == 2 and "piclas_gpu_init" in r.stderr and "CUDA" in r.stderr, r.stderr
    assert not os.path.exists(tmp_path / "res.bin")


@pytest.mark.gpu
def test_cpp_host_and_python_host_agree_bit_for_bit(tmp_path):
    from piclas_b200.particle_step import ParticleStep
    build.build_cuda()
    exe = build.build_host()
    nsteps = 3
    path, mesh, prm, PS, spec, elem, E, dt, ids = _case(tmp_path, nsteps)
    res = str(tmp_path / "res.bin")
    r = subprocess.run([exe, path, res], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    c = casefile.read_arrays(res)
    lost = 0
    with ParticleStep(mesh, prm) as gpu:
        gpu.UploadParticles(PS, spec, elem, IsNewPart=np.ones(len(spec), dtype=np.int32), ids=ids)
        for it in range(nsteps):
            src, ns = gpu.Deposition()
            gpu.SetField(E)
            lost += gpu.PushAndTrack(dt, it)
        d = gpu.DownloadParticles()
        ekin, npart = gpu.KineticEnergy()
    assert c["nLost"][0] == lost
    assert np.array_equal(c["ids"], d["ids"]) and np.array_equal(c["PartState"], d["PartState"])
    assert np.array_equal(c["GlobalElemID"], d["GlobalElemID"]) and np.array_equal(c["PartSpecies"], d["PartSpecies"])
    assert np.array_equal(c["PartSource"].reshape(src.shape), src) and np.array_equal(c["NodeSource"], ns)
    assert np.array_equal(c["Ekin"], ekin) and np.array_equal(c["nPart"], npart)
