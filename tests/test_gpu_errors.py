"""Error behaviour of the C ABI (the replaced Fortran aborts through CALL abort(__STAMP__,...); here every entry point returns
nonzero and piclas_gpu_last_error() carries the message): unsupported configurations are rejected at init, never ignored
(SURVEY.md Appendix A.15), and the call order of the step is enforced."""
import ctypes as C

import numpy as np
import pytest

import cases
from piclas_b200 import hostmesh as hm
from piclas_b200.abi import DEPO_CVW, DEPO_SF
from piclas_b200.particle_step import ParticleStep, PiclasGpuError

pytestmark = pytest.mark.gpu


def _plasma(mesh, n=500):
    PS, spec = cases.uniform_plasma(mesh, n, seed=1, vth_cells=0.3, dt=1e-8)
    return PS, spec, hm.cartesian_locate(mesh, PS[:, :3])


def test_unsupported_configurations_are_rejected_at_init():
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (3, 3, 3), 2)
    with pytest.raises(PiclasGpuError, match="PIC-Deposition-Type"):
        ParticleStep(mesh, cases.electron_params(DepositionType=DEPO_CVW))
    with pytest.raises(PiclasGpuError, match="TimeDiscMethod"):
        ParticleStep(mesh, cases.electron_params(TimeDiscMethod=1))
    with pytest.raises(PiclasGpuError, match="CartesianPeriodic"):
        ParticleStep(mesh, cases.electron_params(CartesianPeriodic=1))
    with pytest.raises(PiclasGpuError, match="cell_volweight_mean requires TrackingMethod=triatracking"):
        ParticleStep(mesh, cases.electron_params(TrackingMethod=hm.REFMAPPING))
    with pytest.raises(PiclasGpuError, match="shape-function deposition needs"):
        ParticleStep(mesh, cases.electron_params(DepositionType=DEPO_SF))          # no FIBGM tables
    wall = hm.box_mesh([0, 0, 0], [1, 1, 1], (3, 3, 3), 2, periodic=(False, True, True), wall_kind=7)
    with pytest.raises(PiclasGpuError, match="TargetBoundCond=7"):
        ParticleStep(wall, cases.electron_params())
    # the context is usable again after a failed init
    with ParticleStep(mesh, cases.electron_params()) as gpu:
        assert gpu.NumParticles() == 0


def test_call_order_and_buffer_checks():
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (3, 3, 3), 2)
    PS, spec, elem = _plasma(mesh)
    with ParticleStep(mesh, cases.electron_params()) as gpu:
        gpu.UploadParticles(PS, spec, elem)
        with pytest.raises(PiclasGpuError, match="no field set"):
            gpu.PushAndTrack(1e-8)
        with pytest.raises(PiclasGpuError):
            gpu.SetField(np.zeros((5, 3, 3, 3, 3)))                                # wrong shape
        n_out = C.c_int64(0)
        small = np.zeros((10, 6))
        rc = gpu.lib.piclas_gpu_download_particles(C.c_int64(10), small.ctypes.data_as(C.POINTER(C.c_double)), None, None, None, None,
                                                   C.byref(n_out))
        assert rc != 0 and b"do not fit" in gpu.lib.piclas_gpu_last_error()
    with ParticleStep(mesh, cases.electron_params(DoDeposition=0)) as gpu:
        gpu.UploadParticles(PS, spec, elem)
        with pytest.raises(PiclasGpuError, match="DoDeposition"):
            gpu.Deposition()


def test_particles_outside_the_rank_are_refused_and_dead_particles_dropped():
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (4, 2, 2), 1)
    off = hm.partition(mesh, 2)
    PS, spec, elem = _plasma(mesh, 400)
    prm = cases.electron_params(nRanks=2, myRank=0)
    with ParticleStep(mesh, prm, offsetElem=int(off[0]), nElems=int(off[1] - off[0])) as gpu:
        with pytest.raises(PiclasGpuError, match="elements of other ranks"):
            gpu.UploadParticles(PS, spec, elem)
    mine = elem <= off[1]
    with ParticleStep(mesh, prm, offsetElem=int(off[0]), nElems=int(off[1] - off[0])) as gpu:
        inside = np.ones(int(mine.sum()), dtype=np.int32)
        inside[::7] = 0                                                            # PDM%ParticleInside = F: not taken over
        gpu.UploadParticles(PS[mine], spec[mine], elem[mine], ParticleInside=inside)
        assert gpu.NumParticles() == int(inside.sum())
        gpu.SetField(np.zeros(gpu._e_shape))
        gpu.PushAndTrack(1e-8)
        # several ranks: the step stays open until the exchange has been finished
        with pytest.raises(PiclasGpuError, match="exchange of the last step is still open"):
            gpu.Deposition()
        cs = C.c_int32(0)
        nsend = (C.c_int64 * 2)()
        sp = C.c_void_p(0)
        assert gpu.lib.piclas_gpu_exchange_info(C.byref(cs), nsend, C.byref(sp)) == 0
        assert nsend[0] == 0 and cs.value == 9                                     # 8 doubles + particle id
        assert gpu.lib.piclas_gpu_exchange_finish(C.c_int64(0)) == 0               # nobody arrives; the emigrants are gone
        assert gpu.NumParticles() == int(inside.sum()) - int(nsend[1])
        gpu.Deposition()


def test_host_input_that_indexes_device_tables_is_range_checked():
    """GlobalElemID / PartSpecies go straight into device table lookups; ELEM_RANK must agree with offsetElem / nElems."""
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (3, 3, 3), 2)
    PS, spec, elem = _plasma(mesh)
    with ParticleStep(mesh, cases.electron_params()) as gpu:
        bad = elem.copy()
        bad[17] = mesh.nElems + 5
        with pytest.raises(PiclasGpuError, match="GlobalElemID outside"):
            gpu.UploadParticles(PS, spec, bad)
        assert gpu.NumParticles() == 0
        bs = spec.copy()
        bs[3] = 2                                                                  # one species configured
        with pytest.raises(PiclasGpuError, match="PartSpecies outside"):
            gpu.UploadParticles(PS, bs, elem)
        bad[17] = 0                                                                # dead slots may carry any element id
        inside = np.ones(len(spec), dtype=np.int32)
        inside[17] = 0
        gpu.UploadParticles(PS, spec, bad, ParticleInside=inside)
        assert gpu.NumParticles() == len(spec) - 1
    off = hm.partition(mesh, 2)
    prm = cases.electron_params(nRanks=2, myRank=0)
    with pytest.raises(PiclasGpuError, match="disagrees with offsetElem"):
        ParticleStep(mesh, prm, offsetElem=int(off[0]), nElems=int(off[1] - off[0]) - 1)
    mesh.ElemInfo[5, 6], mesh.ElemInfo[6, 6] = 1, 0                                # ranks interleaved
    with pytest.raises(PiclasGpuError, match="not ascending|disagrees"):
        ParticleStep(mesh, prm, offsetElem=int(off[0]), nElems=int(off[1] - off[0]))


@pytest.mark.parametrize("arith", [0, 1])
def test_partposref_download_after_cvwm_deposit(arith):
    """with the restructured arithmetic the deposition does not store xi on affine elements; the downloaded PartPosRef
    must still be the reference position of the current particle position (GetPositionInRefElem)."""
    from oracle_lib import Oracle
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (3, 4, 3), 2)
    PS, spec, elem = _plasma(mesh, 2000)
    prm = cases.electron_params(arithmetic=arith)
    orc = Oracle(mesh, prm)
    with ParticleStep(mesh, prm) as gpu:
        gpu.UploadParticles(PS, spec, elem, ids=np.arange(len(spec)))
        gpu.SetField(cases.smooth_field(mesh, amp=1e-4))
        for it in range(2):
            gpu.Deposition()
            d = gpu.DownloadParticles(want_ref=True)
            xi, suc, _ = orc.position_in_ref_elem(d["PartState"][:, :3], d["GlobalElemID"])
            assert suc.all()
            assert np.abs(d["PartPosRef"] - xi).max() <= 1e-12
            gpu.PushAndTrack(1e-8, it)
        d = gpu.DownloadParticles(want_ref=True)                                   # no deposition since the push: still current
        xi, _, _ = orc.position_in_ref_elem(d["PartState"][:, :3], d["GlobalElemID"])
        assert np.abs(d["PartPosRef"] - xi).max() <= 1e-12
    orc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("field,match", [("PartLorentzType", "LorentzType"), ("NoDirichletDeposition", "DoDirichletDeposition"),
                                         ("DoDielectricSurfaceCharge", "DielectricSurfaceCharge")])
def test_unimplemented_reference_options_are_rejected_at_init(field, match):
    """Options of the reference that change the sources / the push and are not implemented abort at init instead of being ignored."""
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (2, 2, 2), 2)
    prm = cases.electron_params(**{field: 1})
    with pytest.raises(PiclasGpuError, match=match):
        ParticleStep(mesh, prm)
