"""The five BASELINE.json configs at test size: same mesh shape, degree, tracking, deposition, species and time step as
the files in the reference tree (SURVEY.md F6 lists what the files actually contain); particle numbers reduced where noted.
Every case runs Deposition + interpolate/push/track for a few steps on the GPU and is compared with the oracle."""
import numpy as np
import pytest

import cases
from oracle_lib import Oracle
from piclas_b200 import hostmesh as hm
from piclas_b200.abi import Params, DEPO_CVWM, DEPO_SF, DEPO_SF_ADAPTIVE, TIMEDISC_BORIS_LEAPFROG
from test_gpu_parity import run_parity
from test_gpu_refmapping import run_ref_parity

pytestmark = pytest.mark.gpu


def plasma_wave_case():
    """tutorials/pic-poisson-plasma-wave with the tutorial's own mesh file (datasets of plasma_wave_mesh.h5 in
    tests/golden/hopr_meshes.npz): 60x1x1 elements on [0,6.2831]x[0,0.2]^2, N=5, refmapping, shape_function_adaptive 1-D x,
    alpha=4, adaptive-DOF=10, 400 electrons (sin_deviation, amplitude 0.01, wave number 2) + 400 ions, dt=5e-10."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hopr_meshes.npz"))
    mesh = hm.from_hopr_arrays(*[g["plasma_wave_" + d] for d in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType", "BCNames")],
                               5, tracking=hm.REFMAPPING)
    Lx = 6.2831
    hm.add_fibgm(mesh, deltas=(Lx, 0.2, 0.2), factor=(60, 1, 1))
    hm.add_refmapping_tables(mesh)
    prm = Params(TrackingMethod=hm.REFMAPPING, DepositionType=DEPO_SF_ADAPTIVE, TimeDiscMethod=TIMEDISC_BORIS_LEAPFROG,
                 ChargeIC=(-cases.QE, cases.QE), MassIC=(cases.ME, 1.672621637e-27), MacroParticleFactor=(5e8, 5e8),
                 carryParticleIDs=1)
    hm.shape_function_adaptive_setup(mesh, prm, 4, dim_sf=1, dim_sf_dir=1, sfDepo3D=True, SFAdaptiveDOF=10)
    n = 400
    # SetParticlePositionSinDeviation (particle_emission_tools.f90): x_i = (i - 0.5) L / n + A sin(k 2 pi x_i / L), y = z = mid
    xs = (np.arange(n) + 0.5) * Lx / n
    xe = xs + 0.01 * np.sin(2.0 * 2.0 * np.pi * xs / Lx)
    PS = np.zeros((2 * n, 6))
    PS[:n, 0] = np.mod(xe, Lx)
    PS[n:, 0] = xs
    PS[:, 1:3] = 0.1
    spec = np.concatenate([np.ones(n), 2 * np.ones(n)]).astype(np.int32)
    return mesh, prm, PS, spec


def test_config1_plasma_wave_tutorial():
    mesh, prm, PS, spec = plasma_wave_case()
    orc = Oracle(mesh, prm)
    elem = orc.locate(PS[:, :3])
    orc.close()
    assert (elem > 0).all()
    E = cases.smooth_field(mesh, 5.0)
    run_ref_parity(mesh, prm, PS, spec, elem, E, 5e-10, nsteps=6)


def test_config2_landau_damping_tutorial():
    """tutorials/pic-poisson-landau-damping: 30x1x1, [0,4pi]x1x1, N=4, refmapping, shape_function 1-D x, r=0.5, alpha=10,
    line deposition, normalised constants (c0=1e8), 40000 electrons (here 8000) + 100 ions, dt=0.1."""
    Lx = 4 * np.pi
    mesh = hm.box_mesh([0, 0, 0], [Lx, 1, 1], (30, 1, 1), 4, tracking=hm.REFMAPPING)
    hm.add_fibgm(mesh, deltas=(12.5663706144, 1., 1.), factor=(30., 1., 1.))
    hm.add_refmapping_tables(mesh)
    prm = Params(TrackingMethod=hm.REFMAPPING, DepositionType=DEPO_SF, TimeDiscMethod=TIMEDISC_BORIS_LEAPFROG,
                 ChargeIC=(-1., 1.), MassIC=(1., 1.e5), MacroParticleFactor=(3.14159265358e-4, 0.125663706144),
                 c2_inv=1.0 / (1e8 * 1e8), carryParticleIDs=1)
    hm.shape_function_setup(mesh, prm, 0.5, 10, dim_sf=1, dim_sf_dir=1, sfDepo3D=False)
    rng = np.random.default_rng(7)
    ne, ni = 8000, 100
    x = np.zeros((ne + ni, 3))
    x[:, 0] = rng.random(ne + ni) * Lx
    x[:, 1:] = rng.random((ne + ni, 2))
    v = np.zeros((ne + ni, 3))
    v[:ne, 0] = rng.normal(0, 1.0, ne)
    PS = np.ascontiguousarray(np.concatenate([x, v], axis=1))
    spec = np.concatenate([np.ones(ne), 2 * np.ones(ni)]).astype(np.int32)
    elem = hm.cartesian_locate(mesh, x)
    E = cases.smooth_field(mesh, 0.05)
    run_ref_parity(mesh, prm, PS, spec, elem, E, 0.1, nsteps=6)


def test_config3_two_stream_instability_tutorial():
    """tutorials/pic-poisson-TSI: 801x1x1, [0,4pi]x0.03^2... N=2, TriaTracking, cell_volweight_mean, two electron beams + ions
    (3 x 100000 in the tutorial, 3 x 15000 here)."""
    Lx = 4 * np.pi
    mesh = hm.box_mesh([0, 0, 0], [Lx, 0.03, 0.03], (801, 1, 1), 2)
    prm = Params(DepositionType=DEPO_CVWM, ChargeIC=(-cases.QE, -cases.QE, cases.QE), MassIC=(cases.ME, cases.ME, 1.672621637e-27),
                 MacroParticleFactor=(1e5, 1e5, 1e5), carryParticleIDs=1)
    rng = np.random.default_rng(8)
    m = 15000
    x = np.zeros((3 * m, 3))
    x[:, 0] = rng.random(3 * m) * Lx
    x[:, 1:] = rng.random((3 * m, 2)) * 0.03
    v = np.zeros((3 * m, 3))
    v[:m, 0] = 1.06e8 * 0.1 + rng.normal(0, 1e5, m)
    v[m:2 * m, 0] = -1.06e8 * 0.1 + rng.normal(0, 1e5, m)
    PS = np.ascontiguousarray(np.concatenate([x, v], axis=1))
    spec = np.repeat([1, 2, 3], m).astype(np.int32)
    elem = hm.cartesian_locate(mesh, x)
    E = cases.smooth_field(mesh, 20.0)
    run_parity(mesh, prm, PS, spec, elem, E, 5e-10, nsteps=6)
