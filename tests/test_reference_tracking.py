"""Per-particle tracking results written by the reference itself (tests/golden/tracking_dsmc_reference.npz, extracted by
tests/golden/make_reference_vectors.py from regressioncheck/NIG_tracking_DSMC/{periodic,ANSA_box}).

Both regression checks restart from a state file, run a collisionless, field-free DSMC time step
(timedisc_TimeStep_DSMC.f90:127-149: LastPartPos = PartState, x += v dt, then PerformTracking) with
TrackingMethod = refmapping, tracing, triatracking, and compare `PartInt` (the particle range of every element, i.e. the
element ownership after tracking) with a committed reference state by h5diff (analyze.ini).  With ChargeIC = 0 the
Leapfrog step (timedisc_TimeStepPoisson.f90:124-181) is exactly that push (Pt = E q/m = 0), so the same files pin this
repo's push + tracking against output of the reference:

* periodic: 5x5x5 box [0,2]x[0,1]^2, three periodic vectors, 1000 particles, 200 steps of 1e-4 (about 13 wraps per particle).
  Velocities must be bit-equal, positions agree to 1e-12, every particle must sit in the element the reference put it in.
* ANSA_box: 1331 hexahedra of an unstructured ANSA mesh (a box turned by 45 degrees about z), specular walls, 2000 particles,
  100 steps of 1e-2 (about 30 wall hits per particle).  `PartInt` must be equal: that is the reference's own acceptance
  test.  The wall nodes of the mesh file carry 1e-9 jitter, so the wall normal (triangle normal in TriaTracking, side
  normal in RefMapping, local bilinear normal in Tracing) differs by 6e-9 between the methods and positions drift by 1e-5
  between them after 30 reflections; the state file comes from one of them, so positions are only compared to 1e-4 here
  (the reference compares none), the speed to 1e-10.
"""
import os

import numpy as np
import pytest

from oracle_lib import Oracle
from piclas_b200 import hostmesh as hm
from piclas_b200.abi import Params, TIMEDISC_LEAPFROG, DEPO_CVWM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "tracking_dsmc_reference.npz")
TRACKINGS = [pytest.param(hm.TRIATRACKING, id="triatracking"), pytest.param(hm.REFMAPPING, id="refmapping")]


def _elem_of_range(PartInt, n):
    """Index (0-based, file order) of the element whose PartInt range holds each particle of the state file."""
    e = np.full(n, -1, dtype=np.int64)
    for k in range(PartInt.shape[1]):
        e[PartInt[0, k]:PartInt[1, k]] = k
    assert (e >= 0).all()
    return e


def _params(tracking, mass, eps):
    # push + tracking only (the DSMC step neither interpolates nor deposits); the deposition type only has to be one the
    # tracking method admits
    return Params(TrackingMethod=tracking, TimeDiscMethod=TIMEDISC_LEAPFROG, ChargeIC=(0.0,), MassIC=(mass,),
                  DoInterpolation=0, DoDeposition=0, DepositionType=DEPO_CVWM if tracking == hm.TRIATRACKING else 0, RefMappingEps=eps,
                  carryParticleIDs=1)


def periodic_case(tracking, fibgm_deltas=(0.4, 0.2, 0.2)):
    """NIG_tracking_DSMC/periodic (hopr.ini: Corner (0,0,0)-(2,1,1), nElems 5,5,5; parameter.ini: ManualTimeStep 1e-4,
    tend 2e-2, RefMappingEps 1e-12).  The mesh file is built by HOPR at test time in the reference and is not in its tree; the
    element each HOPR index stands for is taken from the reference's own localisation of the restart particles (every element
    holds some).  The check's own Part-FIBGMdeltas (2,1,1) put all 125 elements into one background cell; the device kernel
    keeps a sorted list of at most 32 candidates per cell (REF_MAX_BGM, csrc/ref.cuh) and visits fuller cells by repeated
    selection (csrc/select.cuh), so the default here is one cell per element (stored-list path) and the last test of this
    file runs the reference's deltas (selection path)."""
    g = np.load(GOLDEN)
    mesh = hm.box_mesh([0, 0, 0], [2, 1, 1], (5, 5, 5), 1, tracking=tracking)
    if tracking == hm.REFMAPPING:
        hm.add_fibgm(mesh, deltas=fibgm_deltas)
        hm.add_refmapping_tables(mesh, RefMappingEps=1e-12)
        assert (mesh.extra["FIBGM"]["nElems"].max() <= 32) == (fibgm_deltas[0] < 2.0)
    PD0, PD1 = g["periodic_PartData0"], g["periodic_PartData1"]
    n = PD0.shape[0]
    elem0 = hm.cartesian_locate(mesh, PD0[:, :3]).astype(np.int32)
    file_elem0 = _elem_of_range(g["periodic_PartInt0"], n)
    to_ours = np.zeros(mesh.nElems, dtype=np.int32)
    for k in range(mesh.nElems):
        u = np.unique(elem0[file_elem0 == k])
        assert len(u) == 1, "the reference localised the restart particles of one element in several of ours"
        to_ours[k] = u[0]
    assert len(np.unique(to_ours)) == mesh.nElems
    elem1 = to_ours[_elem_of_range(g["periodic_PartInt1"], n)]
    return mesh, _params(tracking, 6e-26, 1e-12), PD0, elem0, PD1, elem1, 1e-4, 200


def ansa_case(tracking):
    """NIG_tracking_DSMC/ANSA_box (parameter.ini: reflective BC_Open, ManualTimeStep 1e-2, tend 1, Part-FIBGMdeltas (1,1,1))
    on the reference's mesh file in its own element order."""
    g = np.load(GOLDEN)
    mesh = hm.from_hopr_arrays(*[g["ansa_mesh_" + d] for d in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs",
                                                                "BCType", "BCNames")],
                               1, part_bc={"BC_Open": hm.BC_REFLECTIVE}, tracking=tracking)
    if tracking == hm.REFMAPPING:
        hm.add_fibgm(mesh, deltas=(1.0, 1.0, 1.0))
        hm.add_refmapping_tables(mesh)
        assert mesh.extra["FIBGM"]["nElems"].max() <= 32          # REF_MAX_BGM of the device kernel
    PD0, PD1 = g["ansa_PartData0"], g["ansa_PartData1"]
    n = PD0.shape[0]
    elem0 = (_elem_of_range(g["ansa_PartInt0"], n) + 1).astype(np.int32)
    elem1 = (_elem_of_range(g["ansa_PartInt1"], n) + 1).astype(np.int32)
    return mesh, _params(tracking, 4.652e-26, 1e-4), PD0, elem0, PD1, elem1, 1e-2, 100


def _match(PS, PD1):
    """Row of the reference's end state for each of our particles (the state file is sorted by element, ours by id)."""
    from scipy.spatial import cKDTree
    sc = np.array([1.0, 1.0, 1.0, 1e3, 1e3, 1e3])
    _, ii = cKDTree(PD1[:, :6] / sc).query(PS / sc)
    assert len(np.unique(ii)) == len(ii), "end states do not pair up one to one"
    return ii


def check_periodic(PS, elem, PD1, elem1, nElems):
    ii = _match(PS, PD1)
    assert np.array_equal(PS[:, 3:], PD1[ii, 3:6]), "velocities must be untouched by periodic tracking"
    err = np.abs(PS[:, :3] - PD1[ii, :3]).max() / np.abs(PD1[:, :3]).max()
    assert err <= 1e-12, err
    assert np.array_equal(elem, elem1[ii]), "element ownership differs from the reference's state file"
    assert np.array_equal(np.bincount(elem, minlength=nElems + 1), np.bincount(elem1, minlength=nElems + 1))
    return err


def check_ansa(PS, elem, PD0, PD1, elem1, nElems):
    # PartInt of the reference = particles per element, in the mesh file's element order: the reference's h5diff criterion
    assert np.array_equal(np.bincount(elem, minlength=nElems + 1), np.bincount(elem1, minlength=nElems + 1)), \
        "PartInt differs from the reference's state file"
    ii = _match(PS, PD1)
    assert np.array_equal(elem, elem1[ii]), "element ownership differs from the reference's state file"
    assert np.abs(PS[:, :3] - PD1[ii, :3]).max() <= 1e-4 and np.abs(PS[:, 3:] - PD1[ii, 3:6]).max() <= 1e-4
    speed0 = np.linalg.norm(PD0[:, 3:6], axis=1)
    assert np.abs(np.linalg.norm(PS[:, 3:], axis=1) - speed0).max() <= 1e-10 * speed0.max()
    assert np.abs(np.linalg.norm(PD1[ii, 3:6], axis=1) - speed0).max() <= 1e-10 * speed0.max()


def run_oracle(mesh, prm, PD0, elem0, dt, nsteps):
    n = PD0.shape[0]
    PS = np.ascontiguousarray(PD0[:, :6])
    spec = PD0[:, 6].astype(np.int32)
    elem = elem0.copy()
    inside = np.ones(n, dtype=np.int32)
    isnew = np.zeros(n, dtype=np.int32)           # restart: no half-step for new particles
    E = np.zeros((mesh.nElems, 2, 2, 2, 3))
    orc = Oracle(mesh, prm)
    xi = None
    if prm.TrackingMethod == hm.REFMAPPING:
        xi, suc, bad = orc.position_in_ref_elem(PS[:, :3], elem)
        assert bad == 0 and np.abs(xi).max() < 1.0
    else:
        ins, _ = orc.inside(PS[:, :3], elem)
        assert ins.all(), "restart particles are not inside the elements PartInt names"
    for _ in range(nsteps):
        nlost, _, _ = orc.push_track(dt, PS, spec, elem, inside, isnew, E, PartPosRef=xi)
        assert nlost == 0
    assert inside.all()
    orc.close()
    return PS, elem


def run_gpu(mesh, prm, PD0, elem0, dt, nsteps):
    from piclas_b200.particle_step import ParticleStep
    n = PD0.shape[0]
    spec = PD0[:, 6].astype(np.int32)
    with ParticleStep(mesh, prm) as gpu:
        gpu.UploadParticles(np.ascontiguousarray(PD0[:, :6]), spec, elem0, IsNewPart=np.zeros(n, dtype=np.int32),
                            ids=np.arange(n, dtype=np.int64))
        gpu.SetField(np.zeros((mesh.nElems, 2, 2, 2, 3)))
        for it in range(nsteps):
            assert gpu.PushAndTrack(dt, it) == 0
        assert gpu.NumParticles() == n
        d = gpu.DownloadParticles()
    o = np.argsort(d["ids"], kind="stable")
    assert np.array_equal(d["ids"][o], np.arange(n))
    return d["PartState"][o], d["GlobalElemID"][o]


# ---- the oracle against the reference's files (CPU) -----------------------------------------------------------------------------
@pytest.mark.parametrize("tracking", TRACKINGS)
def test_oracle_reproduces_the_references_periodic_tracking(tracking):
    mesh, prm, PD0, elem0, PD1, elem1, dt, nsteps = periodic_case(tracking)
    PS, elem = run_oracle(mesh, prm, PD0, elem0, dt, nsteps)
    check_periodic(PS, elem, PD1, elem1, mesh.nElems)


@pytest.mark.parametrize("tracking", TRACKINGS)
def test_oracle_reproduces_the_references_partint_on_the_ansa_box(tracking):
    mesh, prm, PD0, elem0, PD1, elem1, dt, nsteps = ansa_case(tracking)
    PS, elem = run_oracle(mesh, prm, PD0, elem0, dt, nsteps)
    check_ansa(PS, elem, PD0, PD1, elem1, mesh.nElems)


@pytest.mark.skipif(not os.path.isdir("/root/reference/regressioncheck"), reason="reference tree not mounted")
def test_tracking_fixture_is_what_the_reference_files_hold():
    from piclas_b200.h5mini import H5File
    g = np.load(GOLDEN)
    d = "/root/reference/regressioncheck/NIG_tracking_DSMC/"
    for tag, f0, f1 in (("periodic", "periodic/periodic_restart_State_000.0000000000000000.h5",
                         "periodic/periodic_reference_State_000.0200000000000000.h5"),
                        ("ansa", "ANSA_box/tildbox_restart_State_000.0000000000000000.h5",
                         "ANSA_box/tildbox_reference_State_001.0000000000000000.h5")):
        a, b = H5File(d + f0), H5File(d + f1)
        assert np.array_equal(a.read("PartData"), g[tag + "_PartData0"])
        assert np.array_equal(a.read("PartInt"), g[tag + "_PartInt0"])
        assert np.array_equal(b.read("PartData").T, g[tag + "_PartData1"])
        assert np.array_equal(b.read("PartInt"), g[tag + "_PartInt1"])
    me = H5File(d + "ANSA_box/tildbox_mesh.h5")
    for ds in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType"):
        assert np.array_equal(me.read(ds), g["ansa_mesh_" + ds])


# ---- the CUDA path against the same files (through the C ABI) -------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("arith", [0, 1], ids=["reference-order", "restructured"])
@pytest.mark.parametrize("tracking", TRACKINGS)
def test_gpu_reproduces_the_references_periodic_tracking(tracking, arith):
    mesh, prm, PD0, elem0, PD1, elem1, dt, nsteps = periodic_case(tracking)
    prm.arithmetic = arith
    PS, elem = run_gpu(mesh, prm, PD0, elem0, dt, nsteps)
    check_periodic(PS, elem, PD1, elem1, mesh.nElems)


@pytest.mark.gpu
@pytest.mark.parametrize("arith", [0, 1], ids=["reference-order", "restructured"])
@pytest.mark.parametrize("tracking", TRACKINGS)
def test_gpu_reproduces_the_references_partint_on_the_ansa_box(tracking, arith):
    mesh, prm, PD0, elem0, PD1, elem1, dt, nsteps = ansa_case(tracking)
    prm.arithmetic = arith
    PS, elem = run_gpu(mesh, prm, PD0, elem0, dt, nsteps)
    check_ansa(PS, elem, PD0, PD1, elem1, mesh.nElems)


# ---- the reference's own background mesh: all 125 elements in one FIBGM cell ----------------------------------------------------
def test_oracle_periodic_tracking_with_the_references_fibgm_deltas():
    mesh, prm, PD0, elem0, PD1, elem1, dt, nsteps = periodic_case(hm.REFMAPPING, fibgm_deltas=(2.0, 1.0, 1.0))
    PS, elem = run_oracle(mesh, prm, PD0, elem0, dt, nsteps)
    check_periodic(PS, elem, PD1, elem1, mesh.nElems)


def test_selection_order_is_the_stable_insertion_sort():
    """csrc/select.cuh compiled for the host: the repeated selection visits 20000 random lists (ties, skipped entries, up to
    140 entries) in exactly the order of the stable InsertionSort the stored-list path of csrc/ref.cuh uses (utils.f90:52-101)."""
    import subprocess
    import tempfile
    src = r'''
#include "select.cuh"
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>
struct Arr { const double* d; double operator()(int i) const { return d[i]; } };
int main() {
  std::mt19937 rng(5);
  const double skip = -1.7976931348623157e308;
  for (int trial = 0; trial < 20000; ++trial) {
    const int n = 1 + rng() % 140;
    std::vector<double> D(n), S; std::vector<int> P(n);
    for (int i = 0; i < n; ++i) { D[i] = (rng() % 10 == 0) ? skip : (double)(rng() % 12) * 0.25; P[i] = i; }
    S = D;
    for (int i = 1; i < n; ++i) { int j = i - 1; const double tr = S[i]; const int ti = P[i];
      while (j >= 0) { if (S[j] <= tr) break; S[j + 1] = S[j]; P[j + 1] = P[j]; --j; } S[j + 1] = tr; P[j + 1] = ti; }
    std::vector<int> want, got;
    for (int i = 0; i < n; ++i) if (S[i] != skip) want.push_back(P[i]);
    SortedVisit sv; const Arr a{D.data()};
    for (int i = next_in_sorted_order(n, a, skip, sv); i >= 0 && (int)got.size() <= n; i = next_in_sorted_order(n, a, skip, sv)) got.push_back(i);
    if (got != want) { std::printf("MISMATCH %d\n", trial); return 1; }
  }
  // keys that do not order (NaN) must not make the visit endless: at most n entries are returned
  {
    const double nan = std::nan("");
    const double D[5] = {1.0, nan, 0.5, nan, 2.0};
    SortedVisit sv; const Arr a{D};
    int returned = 0;
    for (int i = next_in_sorted_order(5, a, skip, sv); i >= 0; i = next_in_sorted_order(5, a, skip, sv)) if (++returned > 5) break;
    if (returned > 5) { std::printf("ENDLESS\n"); return 1; }
  }
  std::printf("OK\n");
  return 0;
}
'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.cpp"), "w").write(src)
        subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "piclas_b200", "csrc"), "-o", os.path.join(d, "t"),
                        os.path.join(d, "t.cpp")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "OK", out.stdout + out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("arith", [0, 1], ids=["reference-order", "restructured"])
def test_gpu_periodic_tracking_with_the_references_fibgm_deltas(arith):
    """Relocation by repeated selection (more than REF_MAX_BGM candidates in the cell)."""
    mesh, prm, PD0, elem0, PD1, elem1, dt, nsteps = periodic_case(hm.REFMAPPING, fibgm_deltas=(2.0, 1.0, 1.0))
    prm.arithmetic = arith
    PS, elem = run_gpu(mesh, prm, PD0, elem0, dt, nsteps)
    check_periodic(PS, elem, PD1, elem1, mesh.nElems)
