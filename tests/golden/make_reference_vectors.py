#!/usr/bin/env python
"""Extracts golden vectors from files the reference's own regression checks hold (run where /root/reference is mounted):

  regressioncheck/NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean/
      plasma_wave_State_000.00000000000000000_restart.h5   the reference state its h5diff check compares against
                                                           (analyze.ini: DG_Source, DG_Solution): PartData = the 3333
                                                           particles, DG_Source = PartSource(1:4,i,j,k,iElem) as deposited
                                                           by the reference's cell_volweight_mean
      Box_mesh.h5                                          HOPR mesh (element order along the space-filling curve)
  regressioncheck/NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean_save_CVWM/Box_deformed_mesh.h5
                                                           corner nodes of the two-element twisted mesh

  regressioncheck/NIG_tracking_DSMC/periodic/          collisionless, field-free particles in a 3-periodic 5x5x5 box
      periodic_restart_State_000.0000000000000000.h5       PartData / PartInt at t = 0 (1000 particles)
      periodic_reference_State_000.0200000000000000.h5     PartData / PartInt after 200 steps of 1e-4, written by the
                                                           reference (its h5diff check compares PartInt)
  regressioncheck/NIG_tracking_DSMC/ANSA_box/          the same on an unstructured 1331-element box with specular walls
      tildbox_mesh.h5, tildbox_restart_State_000...h5, tildbox_reference_State_001...h5   (2000 particles, 100 steps of 1e-2)

  regressioncheck/NIG_PIC_poisson_Leapfrog/2D_innerBC_dielectric_surface_charge/
      2D_dielectric_innerBC_mesh.h5                        101 hexahedra of several sizes, N = 1
      2Dplasma_test_State_000.00000000000000000.h5         initial state: 791 moving electrons and ions and the DG_Source
                                                           (current and charge density) the reference deposited from them with
                                                           cell_volweight_mean; no surface charge yet (DG_SourceExt = 0)
  regressioncheck/NIG_PIC_poisson_Leapfrog/parallel_plates/PartAnalyzeLeapfrog_ref.csv
                                                           coupled power (kinetic-energy gain per step / dt) of one electron in
                                                           the uniform field of a plate capacitor, all 1500 Leapfrog steps
  regressioncheck/NIG_PIC_maxwell_RK4/single_particle/
      single-particle_mesh.h5, single-particle_reference_State_000.0000000500000000.h5
                                                           one fast electron on a 3x3x3 mesh, N = 3: DG_Source(1:4) deposited by
                                                           the reference with shape_function (r = 0.2, alpha = 4, 3-D) at the last
                                                           Runge-Kutta stage of the run, and the particle's end state
  regressioncheck/WEK_PIC_maxwell/plasma_wave/plasma_wave_restart_State_000.00000030000000000.h5
                                                           3200 electrons and ions at rest on the 60-element mesh of the plasma-wave
                                                           tutorial (same mesh file), N = 5, and the charge density the reference
                                                           deposited from them with the 1-D shape_function (r = 0.15, alpha = 8)
  regressioncheck/NIG_PIC_poisson_plasma_wave/poisson/plasma_wave_restart_State_000.00000000000000000.h5
                                                           the 25 + 25 particles the reference's sin_deviation emission placed
                                                           (initial condition of the plasma-wave configuration)

-> tests/golden/emission_sin_deviation_reference.npz, tests/golden/sf_plasma_wave_reference.npz, tests/golden/sf_single_particle_reference.npz, tests/golden/parallel_plates_pcoupled_reference.npz, tests/golden/plasma_ball_cvwm_reference.npz, tests/golden/hopr_meshes.npz, tests/golden/tracking_dsmc_reference.npz,
   tests/golden/cvwm_current_reference.npz
   (committed; the tests never read /root/reference).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from piclas_b200.h5mini import H5File  # noqa: E402

REF = "/root/reference/regressioncheck/NIG_PIC_Deposition"
TRK = "/root/reference/regressioncheck/NIG_tracking_DSMC"


def current_density_vectors():
    """DG_Source with all four components (moving particles) from the initial state of a PIC-Poisson regression check.  Only
    the initial state qualifies: later state files hold the source deposited at the start of the last step next to the
    particles at its end."""
    d = "/root/reference/regressioncheck/NIG_PIC_poisson_Leapfrog/2D_innerBC_dielectric_surface_charge"
    st = H5File(os.path.join(d, "2Dplasma_test_State_000.00000000000000000.h5"))
    me = H5File(os.path.join(d, "2D_dielectric_innerBC_mesh.h5"))
    out = {"PartData": st.read("PartData"), "PartInt": st.read("PartInt"), "DG_Source": st.read("DG_Source")}
    assert out["PartData"].shape == (791, 7) and out["PartInt"].shape == (101, 2) and out["DG_Source"].shape == (101, 2, 2, 2, 4)
    assert not st.read("DG_SourceExt").any()
    for ds in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType", "BCNames"):
        out["mesh_" + ds] = me.read(ds)
    path = os.path.join(HERE, "cvwm_current_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def coupled_power_vectors():
    ref = np.loadtxt("/root/reference/regressioncheck/NIG_PIC_poisson_Leapfrog/parallel_plates/PartAnalyzeLeapfrog_ref.csv",
                     delimiter=",", skiprows=1)
    assert ref.shape == (1501, 3)
    path = os.path.join(HERE, "parallel_plates_pcoupled_reference.npz")
    np.savez_compressed(path, time=ref[:, 0], PCoupled=ref[:, 1])
    print("wrote", path, os.path.getsize(path), "bytes")


def shape_function_vectors():
    d = "/root/reference/regressioncheck/NIG_PIC_maxwell_RK4/single_particle"
    st = H5File(os.path.join(d, "single-particle_reference_State_000.0000000500000000.h5"))
    me = H5File(os.path.join(d, "single-particle_mesh.h5"))
    out = {"PartData": np.ascontiguousarray(st.read("PartData").T), "DG_Source": st.read("DG_Source")}
    assert out["PartData"].shape == (1, 7) and out["DG_Source"].shape == (27, 4, 4, 4, 4)
    for ds in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType", "BCNames"):
        out["mesh_" + ds] = me.read(ds)
    path = os.path.join(HERE, "sf_single_particle_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def shape_function_1d_vectors():
    d = "/root/reference/regressioncheck/WEK_PIC_maxwell/plasma_wave"
    st = H5File(os.path.join(d, "plasma_wave_restart_State_000.00000030000000000.h5"))
    part, src = st.read("PartData"), st.read("DG_Source")
    assert part.shape == (3200, 7) and src.shape == (60, 6, 6, 6, 4)
    assert not part[:, 3:6].any() and not src[..., :3].any()          # particles at rest: no current density
    me = H5File(os.path.join(d, "plasma_wave_mesh.h5"))               # the tutorial's mesh file, already in hopr_meshes.npz
    tw = H5File("/root/reference/tutorials/pic-poisson-plasma-wave/plasma_wave_mesh.h5")
    for ds in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType"):
        assert np.array_equal(me.read(ds), tw.read(ds))
    path = os.path.join(HERE, "sf_plasma_wave_reference.npz")
    np.savez_compressed(path, PartData=part, PartInt=st.read("PartInt"), DG_Source_charge=np.ascontiguousarray(src[..., 3]))
    print("wrote", path, os.path.getsize(path), "bytes")


def emission_vectors():
    st = H5File("/root/reference/regressioncheck/NIG_PIC_poisson_plasma_wave/poisson/plasma_wave_restart_State_000.00000000000000000.h5")
    part = st.read("PartData")
    assert part.shape == (50, 7)
    path = os.path.join(HERE, "emission_sin_deviation_reference.npz")
    np.savez_compressed(path, PartData=part)
    print("wrote", path, os.path.getsize(path), "bytes")


def tracking_vectors():
    """Particle states the reference wrote before and after pure push + tracking (DSMC time step without collisions:
    timedisc_TimeStep_DSMC.f90:127-149 is x += v dt, then PerformTracking)."""
    out = {}
    for tag, d, r0, r1 in (("periodic", "periodic", "periodic_restart_State_000.0000000000000000.h5",
                            "periodic_reference_State_000.0200000000000000.h5"),
                           ("ansa", "ANSA_box", "tildbox_restart_State_000.0000000000000000.h5",
                            "tildbox_reference_State_001.0000000000000000.h5")):
        a, b = H5File(os.path.join(TRK, d, r0)), H5File(os.path.join(TRK, d, r1))
        pd0, pi0, pd1, pi1 = a.read("PartData"), a.read("PartInt"), b.read("PartData"), b.read("PartInt")
        # the reference state files predate the transposition of the particle arrays (analyze.ini: h5diff_flip = T)
        if pd1.shape[0] == 7:
            pd1 = np.ascontiguousarray(pd1.T)
        assert pd0.shape == pd1.shape and pd0.shape[1] == 7 and pi0.shape == pi1.shape and pi0.shape[0] == 2
        assert pi0[1, -1] == pd0.shape[0] and pi1[1, -1] == pd1.shape[0]
        out[tag + "_PartData0"], out[tag + "_PartInt0"] = pd0, pi0
        out[tag + "_PartData1"], out[tag + "_PartInt1"] = pd1, pi1
    me = H5File(os.path.join(TRK, "ANSA_box", "tildbox_mesh.h5"))
    for ds in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType", "BCNames"):
        out["ansa_mesh_" + ds] = me.read(ds)
    path = os.path.join(HERE, "tracking_dsmc_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    d = os.path.join(REF, "Plasma_Ball_cell_volweight_mean")
    st = H5File(os.path.join(d, "plasma_wave_State_000.00000000000000000_restart.h5"))
    me = H5File(os.path.join(d, "Box_mesh.h5"))
    part = st.read("PartData")                   # (3333, 7): PartState(1:6), species
    src = st.read("DG_Source")                   # (nElems, k, j, i, 4)
    bary = me.read("ElemBarycenters")            # (nElems, 3), HOPR element order
    nodes = me.read("NodeCoords").reshape(-1, 8, 3)
    assert part.shape == (3333, 7) and src.shape == (1000, 2, 2, 2, 4) and bary.shape == (1000, 3)
    assert not src[..., :3].any()                # particles at rest: no current density
    # the elements of this Cartesian HOPR mesh are aligned with the axes (xi = x, eta = y, zeta = z)
    assert np.allclose(nodes[:, 1] - nodes[:, 0], [0.2, 0, 0]) and np.allclose(nodes[:, 2] - nodes[:, 0], [0, 0.2, 0])
    assert np.allclose(nodes[:, 4] - nodes[:, 0], [0, 0, 0.2])
    dm = H5File(os.path.join(REF, "Plasma_Ball_cell_volweight_mean_save_CVWM", "Box_deformed_mesh.h5"))
    dnodes = dm.read("NodeCoords").reshape(-1, 8, 3)     # (2, 8, 3) tensor-ordered corner nodes of the two elements
    # the two HOPR mesh files themselves (datasets hostmesh.from_hopr_arrays consumes)
    meshes = {}
    tw = H5File("/root/reference/tutorials/pic-poisson-plasma-wave/plasma_wave_mesh.h5")
    for tag, f in (("box", me), ("deformed", dm), ("plasma_wave", tw)):
        for ds in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType", "BCNames"):
            meshes[tag + "_" + ds] = f.read(ds)
    mout = os.path.join(HERE, "hopr_meshes.npz")
    np.savez_compressed(mout, **meshes)
    print("wrote", mout, os.path.getsize(mout), "bytes")
    out = os.path.join(HERE, "plasma_ball_cvwm_reference.npz")
    np.savez_compressed(out, PartData=part, DG_Source_charge=np.ascontiguousarray(src[..., 3]), ElemBarycenters=bary,
                        deformed_mesh_NodeCoords=dnodes, PartInt=st.read("PartInt"))
    print("wrote", out, os.path.getsize(out), "bytes")
    tracking_vectors()
    current_density_vectors()
    coupled_power_vectors()
    shape_function_vectors()
    shape_function_1d_vectors()
    emission_vectors()


if __name__ == "__main__":
    main()
