"""Initial emission on the device (piclas_gpu_emit_lattice): positions of SetParticlePositionSinDeviation /
SetParticlePositionCosDistribution (particle_emission_tools.f90:1235-1371) and the element SinglePointToElement assigns
(particle_localization.f90:81-190) against the harness' restatements (cases.sin_deviation is pinned on the reference's own
restart file, tests/test_reference_emission.py) and the oracle's element tests.  Bar: elements bit-exact, positions within
2 ulp of the largest coordinate (the device's sin / cos are not glibc's), y and z (no transcendental) bitwise."""
import numpy as np
import pytest

import cases
from oracle_lib import Oracle
from piclas_b200 import hostmesh as hm
from piclas_b200.particle_step import ParticleStep, PiclasGpuError

pytestmark = pytest.mark.gpu


def _mesh(tracking, nel=(6, 3, 2), lo=(0.0, 0.0, 0.0), hi=(6.2831, 0.9, 0.4), deform=None):
    mesh = hm.box_mesh(list(lo), list(hi), nel, 2, tracking=tracking, deform=deform)
    hm.add_fibgm(mesh)
    if tracking == hm.REFMAPPING:
        hm.add_refmapping_tables(mesh)
    return mesh


def _emit_and_compare(mesh, prm, kind, n3, amp, wn, velo, expected_pos, want_ref=False):
    orc = Oracle(mesh, prm)
    ref = prm.TrackingMethod == hm.REFMAPPING
    el_o = cases.single_point_to_element(mesh, orc, expected_pos, refmapping=ref)
    with ParticleStep(mesh, prm) as gpu:
        n = gpu.EmitLattice(kind, 1, n3, Amplitude=amp, WaveNumber=wn, velocity=velo)
        d = gpu.DownloadParticles(want_ref=want_ref)
    keep = el_o > 0
    assert n == int(keep.sum()) == len(d["ids"])
    o = np.argsort(d["ids"])
    ids = d["ids"][o]
    assert np.array_equal(ids, np.nonzero(keep)[0])                              # id = place in the reference's loop nest
    assert np.array_equal(d["GlobalElemID"][o], el_o[keep]), "element of an emitted particle differs from SinglePointToElement"
    X = d["PartState"][o]
    scale = np.abs(expected_pos).max()
    assert np.abs(X[:, 0] - expected_pos[keep, 0]).max() <= 4.5e-16 * scale
    assert np.array_equal(X[:, 1:3], expected_pos[keep, 1:3])
    assert np.array_equal(X[:, 3:6], np.broadcast_to(np.asarray(velo, dtype=np.float64), (len(X), 3)))
    return d


@pytest.mark.parametrize("amp,wn", [(0.01, 2.0), (0.0, 0.0)])
def test_sin_deviation_triatracking(amp, wn):
    # ny = 6 on 3 elements, nz = 4 on 2: half of the lattice lies exactly on element faces (ConcaveElemSide decides the owner)
    mesh = _mesh(hm.TRIATRACKING)
    prm = cases.electron_params()
    exp = cases.sin_deviation(mesh.xyz_min, mesh.xyz_max, 25, 6, 4, amp, wn)
    _emit_and_compare(mesh, prm, "sin_deviation", (25, 6, 4), amp, wn, (1.0e5, 0.0, -2.0e4), exp)


def test_cos_distribution_triatracking():
    mesh = _mesh(hm.TRIATRACKING)
    prm = cases.electron_params()
    exp = cases.cos_distribution(mesh.xyz_min, mesh.xyz_max, 40, 3, 3, 0.05, 0.5)
    _emit_and_compare(mesh, prm, "cos_distribution", (40, 3, 3), 0.05, 0.5, (0.0, 0.0, 0.0), exp)


def test_sin_deviation_on_a_deformed_mesh():
    """Trilinear elements with non-planar inner sides: the FIBGM cell lists several candidates per point, the order of the
    barycentre distances and the determinant test decide."""
    lo, hi = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    mesh = _mesh(hm.TRIATRACKING, nel=(5, 4, 4), lo=lo, hi=hi, deform=cases.wavy_periodic(0.04, list(lo), list(hi)))
    prm = cases.electron_params()
    exp = cases.sin_deviation(mesh.xyz_min, mesh.xyz_max, 23, 9, 7, 0.02, 3.0)
    _emit_and_compare(mesh, prm, "sin_deviation", (23, 9, 7), 0.02, 3.0, (0.0, 3.0e4, 0.0), exp)


def test_sin_deviation_refmapping():
    from piclas_b200.abi import DEPO_SF
    lo, hi = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    mesh = _mesh(hm.REFMAPPING, nel=(5, 4, 4), lo=lo, hi=hi, deform=cases.wavy(0.04, list(lo), list(hi)))
    prm = cases.electron_params(TrackingMethod=hm.REFMAPPING, DepositionType=DEPO_SF, DoDeposition=0)
    exp = cases.sin_deviation(mesh.xyz_min, mesh.xyz_max, 17, 6, 5, 0.01, 1.0)
    d = _emit_and_compare(mesh, prm, "sin_deviation", (17, 6, 5), 0.01, 1.0, (0.0, 0.0, 0.0), exp, want_ref=True)
    # PartPosRef = GetPositionInRefElem in the element found (particle_localization.f90:73)
    orc = Oracle(mesh, prm)
    xi, suc, _ = orc.position_in_ref_elem(d["PartState"][:, :3], d["GlobalElemID"], force=False)
    assert suc.all() and np.abs(d["PartPosRef"] - xi).max() <= 1e-12


def test_emitted_particles_take_the_new_particle_half_step():
    """IsNewPart = T after the emission (particle_localization.f90:72): the first Boris-Leapfrog step starts with the half step back."""
    mesh = _mesh(hm.TRIATRACKING, nel=(4, 4, 4), hi=(1.0, 1.0, 1.0))
    prm = cases.electron_params()
    exp = cases.sin_deviation(mesh.xyz_min, mesh.xyz_max, 12, 7, 5, 0.02, 1.0)
    orc = Oracle(mesh, prm)
    el = cases.single_point_to_element(mesh, orc, exp)
    assert (el > 0).all()
    E = cases.smooth_field(mesh, amp=1e-2)
    n = len(exp)
    PS = np.concatenate([exp, np.zeros((n, 3))], axis=1)
    spec = np.ones(n, dtype=np.int32)
    dt = 1e-9
    with ParticleStep(mesh, prm) as gpu:
        assert gpu.EmitLattice("sin_deviation", 1, (12, 7, 5), Amplitude=0.02, WaveNumber=1.0) == n
        gpu.SetField(E)
        gpu.PushAndTrack(dt)
        d = gpu.DownloadParticles()
    PSo, elo = PS.copy(), el.copy()
    orc.push_track(dt, PSo, spec, elo, np.ones(n, dtype=np.int32), np.ones(n, dtype=np.int32), E)
    o = np.argsort(d["ids"])
    assert np.array_equal(d["GlobalElemID"][o], elo)
    assert np.abs(d["PartState"][o] - PSo).max() <= 1e-12 * np.abs(PSo).max()


def test_append_and_errors():
    mesh = _mesh(hm.TRIATRACKING)
    prm = cases.electron_params()
    with ParticleStep(mesh, prm) as gpu:
        a = gpu.EmitLattice("sin_deviation", 1, (10, 2, 2), Amplitude=0.01, WaveNumber=2.0)
        b = gpu.EmitLattice("sin_deviation", 1, (5, 2, 2), append=True)
        assert (a, b) == (40, 20) and gpu.NumParticles() == 60
        with pytest.raises(PiclasGpuError, match="species"):
            gpu.EmitLattice("sin_deviation", 3, (5, 2, 2))
        with pytest.raises(PiclasGpuError, match="WaveNumber"):
            gpu.EmitLattice("cos_distribution", 1, (5, 2, 2), Amplitude=0.1, WaveNumber=0.0)
        with pytest.raises(PiclasGpuError, match="SpaceIC"):
            gpu.EmitLattice("cuboid", 1, (5, 2, 2))


def test_device_emission_reproduces_the_references_restart_file():
    """Plasma-wave configuration (BASELINE.json configs[0]): 25 electrons with Amplitude 0.01, WaveNumber 2 and 25 ions with
    Amplitude 0 on [0, 6.2831] x [0, 0.2]^2, 60 x 1 x 1 elements — against PartData of the reference's own restart file
    (tests/golden/emission_sin_deviation_reference.npz, see tests/test_reference_emission.py)."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(os.path.join(root, "tests", "golden", "emission_sin_deviation_reference.npz"))["PartData"]
    mesh = _mesh(hm.TRIATRACKING, nel=(60, 1, 1), hi=(6.2831, 0.2, 0.2))
    prm = cases.electron_params(ChargeIC=(-cases.QE, cases.QE), MassIC=(cases.ME, 1.672621637e-27), MacroParticleFactor=(1.0, 1.0))
    with ParticleStep(mesh, prm) as gpu:
        assert gpu.EmitLattice("sin_deviation", 1, (25, 1, 1), Amplitude=0.01, WaveNumber=2.0) == 25
        assert gpu.EmitLattice("sin_deviation", 2, (25, 1, 1), append=True) == 25
        d = gpu.DownloadParticles()
    for species in (1, 2):
        ref = g[g[:, 6] == species][:, :3]
        ref = ref[np.argsort(ref[:, 0])]
        ours = d["PartState"][d["PartSpecies"] == species][:, :3]
        ours = ours[np.argsort(ours[:, 0])]
        assert ours.shape == ref.shape == (25, 3)
        assert np.abs(ours - ref).max() <= 1e-14
    # the ion at x = 0.3 L lies exactly on the face between elements 18 and 19: both barycentres are equally far, the FIBGM list
    # order and the determinant test (zero is not negative) give it to 18
    orc = Oracle(mesh, prm)
    assert np.array_equal(d["GlobalElemID"], cases.single_point_to_element(mesh, orc, d["PartState"][:, :3]))


def test_every_rank_keeps_the_lattice_points_of_its_own_elements():
    """doHALO = F: each rank calls the emission with the same arguments and accepts the positions inside its own element range
    (one GPU plays the three ranks in turn, as tests/test_gpu_loopback.py does); the union is the single-rank emission."""
    from test_gpu_loopback import LoopRank
    mesh = _mesh(hm.TRIATRACKING, nel=(6, 3, 2))
    n3, amp, wn = (25, 6, 4), 0.01, 2.0
    exp = cases.sin_deviation(mesh.xyz_min, mesh.xyz_max, *n3, amp, wn)
    orc = Oracle(mesh, cases.electron_params())
    el_o = cases.single_point_to_element(mesh, orc, exp)
    world, seen = 3, []
    for r in range(world):
        R = LoopRank(mesh, cases.electron_params(), r, world)
        n = R.step.EmitLattice("sin_deviation", 1, n3, Amplitude=amp, WaveNumber=wn)
        d = R.step.DownloadParticles()
        R.close()
        lo, hi = int(R.off[r]) + 1, int(R.off[r + 1])
        mine = np.nonzero((el_o >= lo) & (el_o <= hi))[0]
        o = np.argsort(d["ids"])
        assert n == len(mine) and np.array_equal(d["ids"][o], mine)
        assert np.array_equal(d["GlobalElemID"][o], el_o[mine])
        seen.append(d["ids"])
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(len(exp)))
