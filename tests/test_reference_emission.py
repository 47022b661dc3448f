"""Initial condition of the plasma-wave configuration (BASELINE.json configs[0]): the positions the reference's `sin_deviation`
emission placed (PartData of regressioncheck/NIG_PIC_poisson_plasma_wave/poisson/plasma_wave_restart_State_000...h5: 25 electrons
with Amplitude 0.01, WaveNumber 2 and 25 ions with Amplitude 0 on [0, 6.2831] x [0, 0.2]^2) against the harness' restatement of
SetParticlePositionSinDeviation (cases.sin_deviation).  Emission itself stays with the host (out of scope); the
restatement only generates test inputs, and this pins it."""
import os

import numpy as np

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sin_deviation_positions_are_the_references():
    g = np.load(os.path.join(ROOT, "tests", "golden", "emission_sin_deviation_reference.npz"))["PartData"]
    assert not g[:, 3:6].any()                                   # velocityDistribution = constant, VeloIC = 0
    for species, amplitude, wavenumber in ((1, 0.01, 2.0), (2, 0.0, 0.0)):
        ref = g[g[:, 6] == species][:, :3]
        ref = ref[np.argsort(ref[:, 0])]                         # the state file is sorted by element, the emission by i
        ours = cases.sin_deviation([0, 0, 0], [6.2831, 0.2, 0.2], 25, 1, 1, amplitude, wavenumber)
        assert ours.shape == ref.shape == (25, 3)
        assert np.abs(ours - ref).max() <= 1e-14                 # 2 ulp of the largest coordinate
    ele = g[g[:, 6] == 1][:, 0]
    ion = g[g[:, 6] == 2][:, 0]
    assert np.abs(np.sort(ele) - np.sort(ion)).max() > 9e-3      # the electrons carry the displacement


def test_emission_fixture_is_what_the_reference_file_holds():
    import pytest
    if not os.path.isdir("/root/reference/regressioncheck"):
        pytest.skip("reference tree not mounted")
    from piclas_b200.h5mini import H5File
    st = H5File("/root/reference/regressioncheck/NIG_PIC_poisson_plasma_wave/poisson/plasma_wave_restart_State_000.00000000000000000.h5")
    g = np.load(os.path.join(ROOT, "tests", "golden", "emission_sin_deviation_reference.npz"))["PartData"]
    assert np.array_equal(st.read("PartData"), g)


def test_cos_distribution_inverts_its_cumulative_distribution():
    """SetParticlePositionCosDistribution (particle_emission_tools.f90:1299-1371): x solves x + a/w sin(w x) = x_uniform to the
    1e-12 the reference's Newton loop stops at; y, z are the cell centres of the lattice."""
    lo, hi = np.array([0.5, -1.0, 2.0]), np.array([6.5, 1.0, 2.5])
    a, w, n3 = 0.3, 1.5, (37, 3, 2)
    X = cases.cos_distribution(lo, hi, *n3, a, w)
    assert X.shape == (37 * 3 * 2, 3)
    xs = X[::6, 0] - lo[0]
    xu = (np.arange(1, 38) - 0.5) * (6.0 / 37)
    assert np.abs(xs + a / w * np.sin(w * xs) - xu).max() <= 1e-12
    assert np.all(np.diff(xs) > 0)
    assert np.array_equal(np.unique(X[:, 1]), (lo[1] + np.arange(1, 4) * (2.0 / 3)) - (2.0 / 3) * 0.5)   # left to right, as written there
    assert np.array_equal(np.unique(X[:, 2]), (lo[2] + np.arange(1, 3) * 0.25) - 0.25 * 0.5)


def test_single_point_to_element_restatement_finds_the_containing_element():
    """The harness restatement of SinglePointToElement (cases.single_point_to_element; FIBGM cell, radius filter, nearest
    barycentre first) against the brute-force search of the oracle on a deformed mesh: same element for points off the faces,
    -1 outside the mesh and for elements of other ranks (doHALO = F)."""
    from oracle_lib import Oracle
    from piclas_b200 import hostmesh as hm
    lo, hi = [0, 0, 0], [1, 1, 1]
    mesh = hm.box_mesh(lo, hi, (5, 4, 3), 2, deform=cases.wavy(0.05, lo, hi))
    hm.add_fibgm(mesh)
    orc = Oracle(mesh, cases.electron_params())
    rng = np.random.default_rng(5)
    X = rng.random((1500, 3))
    el = cases.single_point_to_element(mesh, orc, X)
    brute = orc.locate(X)
    assert (el > 0).all() and np.array_equal(el, brute)
    out = cases.single_point_to_element(mesh, orc, X + np.array([2.0, 0.0, 0.0]))
    assert (out == -1).all()
    part = cases.single_point_to_element(mesh, orc, X, first=21, last=40)
    assert np.array_equal(part, np.where((brute >= 21) & (brute <= 40), brute, -1))
    orc.close()
