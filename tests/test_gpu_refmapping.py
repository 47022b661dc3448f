"""GPU parity for TrackingMethod = refmapping with shape-function deposition — the combination of
the plasma-wave / Landau tutorials and of NIG_PIC_Deposition/Plasma_Ball_Shape-function-*."""
import numpy as np
import pytest

import cases
from oracle_lib import Oracle
from piclas_b200 import hostmesh as hm
from piclas_b200.abi import DEPO_SF, DEPO_SF_CC, TIMEDISC_LEAPFROG

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def run_ref_parity(mesh, prm, PS0, spec, elem0, E, dt, nsteps, deposit=True):
    from piclas_b200.particle_step import ParticleStep
    n = len(spec)
    orc = Oracle(mesh, prm)
    xi0, _, bad = orc.position_in_ref_elem(PS0[:, :3], elem0, force=False)
    PSo, elo, refo = PS0.copy(), elem0.copy(), xi0.copy()
    inside, isnew = np.ones(n, dtype=np.int32), np.ones(n, dtype=np.int32)
    with ParticleStep(mesh, prm) as gpu:
        gpu.UploadParticles(PS0, spec, elem0, IsNewPart=isnew, ids=np.arange(n))     # PartPosRef computed on the device
        gpu.SetField(E)
        for it in range(nsteps):
            if deposit:
                PSr, _ = orc.deposit(PSo, spec, elo, inside)
                PSg, _ = gpu.Deposition()
                for c in range(4):
                    sc = np.abs(PSr[..., c]).max()
                    assert np.abs(PSg[..., c] - PSr[..., c]).max() <= RTOL * sc, (it, c)
            nlo, _, _ = orc.push_track(dt, PSo, spec, elo, inside, isnew, E, PartPosRef=refo)
            nlg = gpu.PushAndTrack(dt, it)
            d = gpu.DownloadParticles(want_ref=True)
            o = np.argsort(d["ids"])
            alive = np.nonzero(inside)[0]
            assert nlg == nlo
            assert np.array_equal(d["ids"][o], alive)
            assert np.array_equal(d["GlobalElemID"][o], elo[alive]), "element ownership differs (step %d)" % it
            assert np.abs(d["PartState"][o] - PSo[alive]).max() <= RTOL * np.abs(PSo).max()
            assert np.abs(d["PartPosRef"][o] - refo[alive]).max() <= RTOL
    orc.close()


@pytest.mark.parametrize("shape", [(6, 5, 4), (4, 4, 1), (30, 1, 1)])
@pytest.mark.parametrize("arith", [0, 1])
def test_refmapping_periodic_boxes(shape, arith):
    mesh = hm.box_mesh([0, 0, 0], [2, 1, 1], shape, 3, tracking=hm.REFMAPPING)
    hm.add_fibgm(mesh)
    hm.add_refmapping_tables(mesh)
    prm = cases.electron_params(TrackingMethod=hm.REFMAPPING, DepositionType=DEPO_SF_CC, MacroParticleFactor=(1e9,), arithmetic=arith)
    hm.shape_function_setup(mesh, prm, 0.3, 2, dim_sf=1, dim_sf_dir=1, sfDepo3D=False)
    dt = 1e-8
    PS, spec = cases.uniform_plasma(mesh, 12000, seed=31, vth_cells=0.5, dt=dt)
    elem = hm.cartesian_locate(mesh, PS[:, :3])
    E = cases.smooth_field(mesh, 1e-4)
    run_ref_parity(mesh, prm, PS, spec, elem, E, dt, nsteps=5)


def test_refmapping_halo_limited_bc_lists_and_open_walls():
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (8, 8, 8), 2, tracking=hm.REFMAPPING, periodic=(True, False, True))
    hm.add_fibgm(mesh)
    hm.add_refmapping_tables(mesh, bc_halo_eps=0.15)            # inner elements take the pure Newton path
    assert (mesh.extra["ElemToBCSides"][:, 0] <= 0).any()
    prm = cases.electron_params(TrackingMethod=hm.REFMAPPING, DepositionType=DEPO_SF, TimeDiscMethod=TIMEDISC_LEAPFROG,
                                DoDeposition=0)
    dt = 1e-8
    PS, spec = cases.uniform_plasma(mesh, 15000, seed=32, vth_cells=0.6, dt=dt)
    elem = hm.cartesian_locate(mesh, PS[:, :3])
    E = cases.smooth_field(mesh, 1e-4)
    run_ref_parity(mesh, prm, PS, spec, elem, E, dt, nsteps=5, deposit=False)


@pytest.mark.parametrize("arith", [0, 1])
def test_refmapping_deformed_mesh(arith):
    """Unstructured-like mesh: trilinear elements with non-planar inner sides (planar box boundary), RefMapping tracking through
    the Newton mapping and the FIBGM, shape-function deposition — the closest stand-in for the NIG_PIC_Deposition meshes."""
    lo, hi = [0, 0, 0], [1, 1, 1]
    mesh = hm.box_mesh(lo, hi, (5, 4, 4), 2, tracking=hm.REFMAPPING, deform=cases.wavy(0.05, lo, hi))
    hm.add_fibgm(mesh)
    hm.add_refmapping_tables(mesh)
    prm = cases.electron_params(TrackingMethod=hm.REFMAPPING, DepositionType=DEPO_SF, MacroParticleFactor=(1e9,), arithmetic=arith)
    hm.shape_function_setup(mesh, prm, 0.25, 2, dim_sf=3)
    dt = 1e-8
    PS, spec = cases.uniform_plasma(mesh, 8000, seed=33, vth_cells=0.4, dt=dt)
    orc = Oracle(mesh, prm)
    elem = orc.locate(PS[:, :3])
    orc.close()
    assert (elem > 0).all()
    E = cases.smooth_field(mesh, 1e-4)
    run_ref_parity(mesh, prm, PS, spec, elem, E, dt, nsteps=4)


@pytest.mark.parametrize("arith", [0, 1])
def test_refmapping_bilinear_and_nonrect_bc_sides(arith):
    """Periodic mesh with a deformed boundary: BILINEAR x-faces and PLANAR_NONRECT y-/z-faces are crossed through
    ComputeBiLinearIntersection with the bilinear side normal; the fastest particles take the LocateParticleInElement fallback
    (and come back with IsNewPart set).  Shape-function deposition on the same mesh."""
    lo, hi = [0, 0, 0], [1, 1, 1]
    mesh = hm.box_mesh(lo, hi, (4, 4, 3), 2, tracking=hm.REFMAPPING, deform=cases.wavy_periodic(0.04, lo, hi))
    hm.add_fibgm(mesh)
    hm.add_refmapping_tables(mesh)
    prm = cases.electron_params(TrackingMethod=hm.REFMAPPING, DepositionType=DEPO_SF, MacroParticleFactor=(1e9,), arithmetic=arith)
    hm.shape_function_setup(mesh, prm, 0.25, 2, dim_sf=3)
    n, dt = 6000, 1e-8
    rng = np.random.default_rng(2)
    el = rng.integers(1, mesh.nElems + 1, n).astype(np.int32)
    xi = rng.uniform(-0.999, 0.999, (n, 3))
    w = lambda t: np.stack([(1 - t) / 2, (1 + t) / 2], -1)
    W = np.einsum("nk,nj,ni->nkji", w(xi[:, 2]), w(xi[:, 1]), w(xi[:, 0]))
    x = np.einsum("nkji,nkjix->nx", W, mesh.XCL_NGeo[el - 1])
    PS = np.ascontiguousarray(np.concatenate([x, rng.normal(0, 0.25 / dt, (n, 3))], axis=1))
    spec = np.ones(n, dtype=np.int32)
    E = cases.smooth_field(mesh, 1e-4)
    run_ref_parity(mesh, prm, PS, spec, el, E, dt, nsteps=6)


@pytest.mark.parametrize("arith", [0, 1])
def test_refmapping_reflective_walls(arith):
    """Specular walls with RefMapping (GetBoundaryInteraction case 2 -> PerfectReflection in ParticleBCTracking): mirrored
    velocities, several wall hits per step, one periodic direction."""
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (4, 3, 5), 2, tracking=hm.REFMAPPING, periodic=(False, True, False),
                       wall_kind=hm.BC_REFLECTIVE)
    hm.add_fibgm(mesh)
    hm.add_refmapping_tables(mesh)
    prm = cases.electron_params(TrackingMethod=hm.REFMAPPING, DepositionType=DEPO_SF, DoDeposition=0, arithmetic=arith)
    dt = 1e-8
    PS, spec = cases.uniform_plasma(mesh, 12000, seed=43, vth_cells=0.8, dt=dt)
    elem = hm.cartesian_locate(mesh, PS[:, :3])
    E = cases.smooth_field(mesh, 2e-4)
    run_ref_parity(mesh, prm, PS, spec, elem, E, dt, nsteps=5, deposit=False)
