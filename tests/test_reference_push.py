"""Leapfrog push in a uniform field against a time series written by the reference
(tests/golden/parallel_plates_pcoupled_reference.npz = column PCoupled of
regressioncheck/NIG_PIC_poisson_Leapfrog/parallel_plates/PartAnalyzeLeapfrog_ref.csv).

The check puts one electron at rest at (0.1, 0.5, 0.5) between two plates at 0 V (x = 0) and 1000 V (x = 1) on a 5x1x1 mesh
(N = 1, refmapping, periodic in y and z) and runs 1500 Leapfrog steps of 4.58e-11 s (timedisc_TimeStepPoisson.f90:124-203).
`PCoupled` is the kinetic energy the particle gained in the step divided by dt (CalcCoupledPowerPart,
particle_analyze_tools.f90:3540-3580).  The potential is linear, so E = (-1000, 0, 0) V/m at every degree of freedom (the
electron's own field is 1e-9 V/m), and the Leapfrog series is

    P_1 = m a^2 dt / 8,   P_n = (n - 1) m a^2 dt  (n >= 2),   a = q E / m,

i.e. v_n = (n - 1/2) a dt: the half step back of a new particle (`IsNewPart`, `:145-156`) followed by full steps.

What the file holds, and how it is used:
* Row 1 is simulation output of the reference (ChargeIC = 1.60217653e-19 of parameter.ini): it equals m a^2 dt / 8 to
  1.4e-11 (the tolerance of its CG solver).  This repo's first step must match it to 1e-10; that pins the half step back
  for new particles and the interpolation of the field against output of the reference.
* Rows 2..1500 are not simulation output: they equal E^2 q'^2 / m * t_(n-1) to 3e-14 with q' = 1.602176634e-19, which is the
  analytical solution the check's readme.md describes, written by the `P_anlay` test hook quoted there.  They are the
  reference's *known answer* for this case (its tolerance: 1e-2); the Leapfrog series above coincides with that
  expression, so this repo must match them up to the charge constant, (q'/q)^2 - 1 = 1.298e-7, and the closed form with
  its own charge to 1e-12.  (A present-day run of the reference deviates from these rows by a few 1e-3 once
  |v| > 1e6 m/s, because CalcEkinPart switches to (gamma - 1) m c^2 there, particle_analyze_pure.f90:66-76; the kinetic
  energy here is the classical one throughout, as in the analytical solution.)

The run also crosses two element faces under RefMapping (the electron ends at x = 0.515 in element 3).
"""
import os

import numpy as np
import pytest

from oracle_lib import Oracle
from piclas_b200 import hostmesh as hm
from piclas_b200.abi import Params, TIMEDISC_LEAPFROG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "parallel_plates_pcoupled_reference.npz")
Q, M, DT, NSTEPS = -1.60217653E-19, 9.1093826E-31, 4.58E-11, 1500
Q_FILE = 1.602176634e-19            # the constant of the analytical rows (globals_vars.f90:54)


def case():
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (5, 1, 1), 1, periodic=(False, True, True), wall_kind=hm.BC_REFLECTIVE,
                       tracking=hm.REFMAPPING)
    hm.add_fibgm(mesh, deltas=(1.0, 1.0, 1.0))
    hm.add_refmapping_tables(mesh)
    prm = Params(TrackingMethod=hm.REFMAPPING, TimeDiscMethod=TIMEDISC_LEAPFROG, ChargeIC=(Q,), MassIC=(M,),
                 DoDeposition=0, DepositionType=0)
    PS = np.array([[0.1, 0.5, 0.5, 0.0, 0.0, 0.0]])
    E = np.zeros((mesh.nElems, 2, 2, 2, 3))
    E[..., 0] = -1000.0
    return mesh, prm, PS, E, np.load(GOLDEN)["PCoupled"]


def check(ekin, ref):
    """ekin[n] = classical kinetic energy after step n (ekin[0] = 0)."""
    P = np.diff(ekin) / DT
    assert ref.shape == (NSTEPS + 1,) and ref[0] == 0.0
    assert abs(P[0] / ref[1] - 1.0) <= 1e-10                                 # the reference's simulated first step
    base = (Q * 1000.0) ** 2 / M * DT
    n1 = np.arange(1, NSTEPS)
    assert np.abs(P[1:] / (base * n1) - 1.0).max() <= 1e-12                  # closed form of the Leapfrog series
    ratio = (Q_FILE / Q) ** 2
    assert np.abs(ref[2:] / (base * ratio * n1) - 1.0).max() <= 1e-13        # the file's rows are that form with q'
    assert np.abs(P[1:] * ratio / ref[2:] - 1.0).max() <= 1e-12              # ... so we match them up to the constant
    assert np.abs(P[1:] / ref[2:] - 1.0).max() <= 2e-7                       # and as they stand to 1.3e-7


def test_oracle_reproduces_the_references_coupled_power_series():
    mesh, prm, PS, E, ref = case()
    orc = Oracle(mesh, prm)
    spec = np.ones(1, dtype=np.int32)
    elem = orc.locate(PS[:, :3]).astype(np.int32)
    xi, _, bad = orc.position_in_ref_elem(PS[:, :3], elem)
    assert elem[0] == 1 and bad == 0
    inside, isnew = np.ones(1, dtype=np.int32), np.ones(1, dtype=np.int32)
    ekin = [0.0]
    for _ in range(NSTEPS):
        nlost, _, _ = orc.push_track(DT, PS, spec, elem, inside, isnew, E, PartPosRef=xi)
        assert nlost == 0
        ekin.append(0.5 * M * (PS[0, 3:] ** 2).sum())
    orc.close()
    assert elem[0] == 3 and abs(PS[0, 0] - 0.515054) < 1e-6 and PS[0, 1] == 0.5 and PS[0, 2] == 0.5
    check(np.array(ekin), ref)


@pytest.mark.skipif(not os.path.isdir("/root/reference/regressioncheck"), reason="reference tree not mounted")
def test_push_fixture_is_what_the_reference_file_holds():
    ref = np.loadtxt("/root/reference/regressioncheck/NIG_PIC_poisson_Leapfrog/parallel_plates/PartAnalyzeLeapfrog_ref.csv",
                     delimiter=",", skiprows=1)
    g = np.load(GOLDEN)
    assert np.array_equal(ref[:, 0], g["time"]) and np.array_equal(ref[:, 1], g["PCoupled"])


@pytest.mark.gpu
@pytest.mark.parametrize("arith", [0, 1], ids=["reference-order", "restructured"])
def test_gpu_reproduces_the_references_coupled_power_series(arith):
    """The same through the C ABI.  piclas_gpu_kinetic_energy (CalcKineticEnergy on the device, with the reference's switch
    to the relativistic form above 1e6 m/s) must agree with the classical energy while the electron is slower than that."""
    from piclas_b200.particle_step import ParticleStep
    mesh, prm, PS, E, ref = case()
    prm.arithmetic = arith
    ekin = [0.0]
    with ParticleStep(mesh, prm) as gpu:
        gpu.UploadParticles(PS, np.ones(1, dtype=np.int32), np.ones(1, dtype=np.int32), IsNewPart=np.ones(1, dtype=np.int32))
        gpu.SetField(E)
        for it in range(NSTEPS):
            assert gpu.PushAndTrack(DT, it) == 0
            d = gpu.DownloadParticles()
            v2 = (d["PartState"][0, 3:] ** 2).sum()
            ekin.append(0.5 * M * v2)
            if it in (0, 50, 100):                        # |v| = 4e3, 4.1e5, 8.1e5 m/s
                e, n = gpu.KineticEnergy()
                assert n[0] == 1 and abs(e[0] - ekin[-1]) <= 1e-14 * ekin[-1]
    assert d["GlobalElemID"][0] == 3 and abs(d["PartState"][0, 0] - 0.515054) < 1e-6
    check(np.array(ekin), ref)
