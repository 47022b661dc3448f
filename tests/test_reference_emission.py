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
