"""The bodies of the -m gpu tests that compare with files written by the reference, run on the CPU with the oracle standing in
for the device (tests/oracle_step.py).  This does not test the CUDA path — the -m gpu run does — it keeps the test code itself
honest where no GPU exists: a wrong argument, shape, tolerance or bookkeeping step shows up here first."""
import pytest

import piclas_b200.particle_step as ps
from oracle_step import OracleStep
from piclas_b200 import hostmesh as hm

import test_reference_deposition as t_dep
import test_reference_push as t_push
import test_reference_shapefunction as t_sf
import test_reference_tracking as t_trk
import test_zz_gpu_properties as t_prop


@pytest.fixture
def oracle_device(monkeypatch):
    monkeypatch.setattr(ps, "ParticleStep", OracleStep)


def test_deposition_and_shape_function_bodies(oracle_device):
    t_dep.test_gpu_reproduces_the_references_current_and_charge_density(0)
    t_sf.test_gpu_reproduces_the_references_shape_function_source_per_dof(0)
    t_sf.test_gpu_reproduces_the_references_1d_shape_function_charge_density(0)


def test_push_body(oracle_device):
    t_push.test_gpu_reproduces_the_references_coupled_power_series(0)


@pytest.mark.parametrize("tracking", [hm.TRIATRACKING, hm.REFMAPPING], ids=["triatracking", "refmapping"])
def test_tracking_bodies(oracle_device, tracking):
    t_trk.test_gpu_reproduces_the_references_periodic_tracking(tracking, 0)
    t_trk.test_gpu_reproduces_the_references_partint_on_the_ansa_box(tracking, 0)


def test_big_cell_and_property_bodies(oracle_device, monkeypatch):
    t_trk.test_gpu_periodic_tracking_with_the_references_fibgm_deltas(0)
    monkeypatch.setattr(t_prop, "NE", 10)
    monkeypatch.setattr(t_prop, "NPART", 200000)
    t_prop.test_conservation_ownership_layout_and_idempotence(0)


def test_smoke_body(oracle_device, capsys):
    import __graft_entry__ as entry
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out
