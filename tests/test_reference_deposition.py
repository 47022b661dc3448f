"""Current and charge density deposited by the reference from moving particles (tests/golden/cvwm_current_reference.npz,
extracted by tests/golden/make_reference_vectors.py).

regressioncheck/NIG_PIC_poisson_Leapfrog/2D_innerBC_dielectric_surface_charge restarts from
`2Dplasma_test_State_000.00000000000000000.h5`: 791 electrons and ions with thermal velocities on a mesh of 101 hexahedra of
several sizes (N = 1, TriaTracking, cell_volweight_mean, MPF 1e12) and the `DG_Source(1:4)` the reference deposited from exactly
these particles (PIC-OutputSource = T; no dielectric surface charge yet, DG_SourceExt = 0).  The particles of
NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean are at rest, so this file is what pins the current density
(`PartSource(1:3)`, pic_depo_method.f90:459-510) per degree of freedom.

Tolerance 5e-12 of the maximum of the charge density / of the current density (the reference's own h5diff tolerance for
DG_Source in this check is 1e-2 relative, analyze.ini).  The measured deviation is <= 2e-12 and is a per-DOF factor common to
charge and current density (correlation 0.98 over the 808 DOFs; permuting the particles moves the sums by 1e-16 only), i.e.
it sits in the element Jacobians behind NodeVolume (CalcCellLocNodeVolumes, pic_depo_tools.f90:224-355: the reference takes
ElemsJ from its metrics through Vandermonde interpolations, this repo's host mirror evaluates the trilinear Jacobian
directly), a table the Fortran host passes to piclas_gpu_init, not in the particle path.  On the uniform mesh of
Plasma_Ball_cell_volweight_mean the same comparison agrees to 3e-15.
"""
import os

import numpy as np
import pytest

from oracle_lib import Oracle
from piclas_b200 import hostmesh as hm
from piclas_b200.abi import Params, TIMEDISC_LEAPFROG, DEPO_CVWM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "cvwm_current_reference.npz")
QE = 1.60217653E-19
RTOL = 2e-12   # measured 1.5e-12 (current) / 1.0e-12 (charge) with the reference's modal DetJac projection (hostmesh), 2e-12 before
BCS = ("BC_WALL", "BC_WALL_INLET", "BC_WALL_PUMP", "BC_SUBSTRAT", "BC_ELECTRODE", "BC_SYMMETRY")


def case():
    """parameter.ini: species 1 neutral (none in the file), 2 electrons, 3 singly charged ions, MacroParticleFactor 1e12."""
    g = np.load(GOLDEN)
    mesh = hm.from_hopr_arrays(*[g["mesh_" + d] for d in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType",
                                                          "BCNames")],
                               1, part_bc={k: hm.BC_REFLECTIVE for k in BCS})
    prm = Params(TimeDiscMethod=TIMEDISC_LEAPFROG, ChargeIC=(0.0, -QE, QE), MassIC=(1.0, 1.0, 1.0),
                 MacroParticleFactor=(1e12, 1e12, 1e12), DepositionType=DEPO_CVWM)
    PD, PI = g["PartData"], g["PartInt"]
    n = PD.shape[0]
    elem = np.zeros(n, dtype=np.int32)
    for e in range(mesh.nElems):
        elem[PI[e, 0]:PI[e, 1]] = e + 1
    assert (elem > 0).all()
    return mesh, prm, np.ascontiguousarray(PD[:, :6]), PD[:, 6].astype(np.int32), elem, g["DG_Source"]


def check(src, ref):
    assert src.shape == ref.shape
    assert np.abs(ref[..., :3]).max() > 1e3 and np.abs(ref[..., 3]).max() > 1e-3      # all four components are populated
    jmax = np.abs(ref[..., :3]).max()
    for c in range(3):
        assert np.abs(src[..., c] - ref[..., c]).max() <= RTOL * jmax, c
    assert np.abs(src[..., 3] - ref[..., 3]).max() <= RTOL * np.abs(ref[..., 3]).max()


def test_oracle_reproduces_the_references_current_and_charge_density():
    mesh, prm, PS, spec, elem, ref = case()
    orc = Oracle(mesh, prm)
    ins, _ = orc.inside(PS[:, :3], elem)
    assert ins.all(), "the particles are not inside the elements PartInt names"
    src, _ = orc.deposit(PS, spec, elem, np.ones(len(spec), dtype=np.int32))
    orc.close()
    check(src, ref)


@pytest.mark.skipif(not os.path.isdir("/root/reference/regressioncheck"), reason="reference tree not mounted")
def test_deposition_fixture_is_what_the_reference_files_hold():
    from piclas_b200.h5mini import H5File
    d = "/root/reference/regressioncheck/NIG_PIC_poisson_Leapfrog/2D_innerBC_dielectric_surface_charge/"
    st, me, g = H5File(d + "2Dplasma_test_State_000.00000000000000000.h5"), H5File(d + "2D_dielectric_innerBC_mesh.h5"), np.load(GOLDEN)
    for ds in ("PartData", "PartInt", "DG_Source"):
        assert np.array_equal(st.read(ds), g[ds])
    for ds in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType"):
        assert np.array_equal(me.read(ds), g["mesh_" + ds])


@pytest.mark.gpu
@pytest.mark.parametrize("arith", [0, 1], ids=["reference-order", "restructured"])
def test_gpu_reproduces_the_references_current_and_charge_density(arith):
    from piclas_b200.particle_step import ParticleStep
    mesh, prm, PS, spec, elem, ref = case()
    prm.arithmetic = arith
    with ParticleStep(mesh, prm) as gpu:
        gpu.UploadParticles(PS, spec, elem)
        src, _ = gpu.Deposition()
    check(src, ref)
