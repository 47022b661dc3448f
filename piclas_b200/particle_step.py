"""Host-side mirror of the reference's particle-step interface, on top of the C ABI (include/piclas_gpu.h).

The reference exposes this path as argument-less subroutines on module globals, called from
`TimeStepPoissonByBorisLeapfrog` (reference src/timedisc/timedisc_TimeStepPoissonByBorisLeapfrog.f90:93-279):

    Deposition()                                       :93    -> ParticleStep.Deposition()
    HDG(time,iter)                                     :99    -> stays on the host (field_solver callback)
    LastPartPos=..; InterpolateFieldToParticle(); push :109-198 \\
    PerformTracking()                                  :211     > ParticleStep.PushAndTrack(dt)
    MPI particle exchange; UpdateNextFreePosition()    :213-270 /

Everything here is plumbing: numpy <-> host pointers.  All arithmetic happens in libpiclas_gpu.so (CUDA, sm_100a);
there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import numpy as np

from .abi import Marshalled, Params, DEPO_CVWM, c_f64p, c_i32p, c_i64p
from .hostmesh import ParticleMesh
from . import lib as _lib


class PiclasGpuError(RuntimeError):
    """Raised where the Fortran glue would CALL Abort(__STAMP__, msg) (globals/globals.f90:322-397)."""


def _f(a):
    return a.ctypes.data_as(c_f64p) if a is not None else C.cast(None, c_f64p)


def _i(a):
    return a.ctypes.data_as(c_i32p) if a is not None else C.cast(None, c_i32p)


def _l(a):
    return a.ctypes.data_as(c_i64p) if a is not None else C.cast(None, c_i64p)


class ParticleStep:
    """One rank's device-resident particle population + the operators of the particle step."""

    def __init__(self, mesh: ParticleMesh, params: Params, offsetElem: int = 0, nElems: int | None = None):
        self.lib = _lib.load()
        self.mesh = mesh
        self.params = params
        self.offsetElem = offsetElem
        self.nElems = mesh.nElems if nElems is None else nElems
        self._mar = Marshalled(mesh, params, offsetElem=offsetElem, nElems=self.nElems)
        self._check(self.lib.piclas_gpu_init(C.byref(self._mar.mesh), C.byref(self._mar.params)))
        n1 = mesh.N + 1
        self._ps_shape = (self.nElems, n1, n1, n1, 4)
        self._e_shape = (self.nElems, n1, n1, n1, 3)
        self._open = True

    # ------------------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise PiclasGpuError(self.lib.piclas_gpu_last_error().decode())

    def close(self):
        if getattr(self, "_open", False):
            self.lib.piclas_gpu_finalize()
            self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------------------------------------------
    def UploadParticles(self, PartState, PartSpecies, GlobalElemID, ParticleInside=None, IsNewPart=None,
                        PartPosRef=None, ids=None, append=False):
        """PartState(1:6,1:n) etc. as numpy arrays with the Fortran index order reversed ([n,6])."""
        PartState = np.ascontiguousarray(PartState, dtype=np.float64)
        n = PartState.shape[0]
        spec = np.ascontiguousarray(PartSpecies, dtype=np.int32)
        elem = np.ascontiguousarray(GlobalElemID, dtype=np.int32)
        ins = None if ParticleInside is None else np.ascontiguousarray(ParticleInside, dtype=np.int32)
        new = None if IsNewPart is None else np.ascontiguousarray(IsNewPart, dtype=np.int32)
        ref = None if PartPosRef is None else np.ascontiguousarray(PartPosRef, dtype=np.float64)
        idv = None if ids is None else np.ascontiguousarray(ids, dtype=np.int64)
        self._check(self.lib.piclas_gpu_upload_particles(C.c_int64(n), _f(PartState), _i(spec), _i(elem), _i(ins), _i(new),
                                                         _f(ref), _l(idv), C.c_int32(1 if append else 0)))

    def NumParticles(self) -> int:
        return int(self.lib.piclas_gpu_num_particles())

    def DownloadParticles(self, want_ref=False):
        n = self.NumParticles()
        PS = np.zeros((n, 6))
        spec = np.zeros(n, dtype=np.int32)
        elem = np.zeros(n, dtype=np.int32)
        ref = np.zeros((n, 3)) if want_ref else None
        ids = np.zeros(n, dtype=np.int64) if self.params.carryParticleIDs else None
        nout = C.c_int64(0)
        self._check(self.lib.piclas_gpu_download_particles(C.c_int64(n), _f(PS), _i(spec), _i(elem), _f(ref), _l(ids),
                                                           C.byref(nout)))
        return dict(PartState=PS, PartSpecies=spec, GlobalElemID=elem, PartPosRef=ref, ids=ids)

    def FillParticleData(self, offsetnPart=0):
        """PartInt / PartData of WriteParticleToHDF5 (io_hdf5/hdf5_output_particle.f90:1595-1670, FillParticleData):
        PartInt[iElem] = (first, last) offsets of the element's particles, PartData[iPart] = (PartState(1:6), species).  The
        device keeps the particles sorted by element, so the downloaded order already is the file order."""
        d = self.DownloadParticles()
        loc = d["GlobalElemID"].astype(np.int64) - self.offsetElem - 1
        if loc.size and (np.any(np.diff(loc) < 0) or loc.min() < 0 or loc.max() >= self.nElems):
            raise PiclasGpuError("FillParticleData: particles are not sorted by local element")
        cnt = np.bincount(loc, minlength=self.nElems).astype(np.int64)
        last = offsetnPart + np.cumsum(cnt)
        PartInt = np.stack([last - cnt, last], axis=1)
        PartData = np.concatenate([d["PartState"], d["PartSpecies"].astype(np.float64)[:, None]], axis=1)
        return PartInt, np.ascontiguousarray(PartData)

    # ------------------------------------------------------------------------------------------------------
    def Deposition(self, want_partsource=True, want_nodesource=True, out_partsource=None, out_nodesource=None):
        """CALL Deposition() (pic_depo.f90:944-1018): returns (PartSource[nElems,k,j,i,4], NodeSource[nNodes,4])."""
        PS = out_partsource if out_partsource is not None else (np.empty(self._ps_shape) if want_partsource else None)
        if self.params.DepositionType != DEPO_CVWM:
            want_nodesource, out_nodesource = False, None      # NodeSource exists for cell_volweight_mean only
        NS = out_nodesource if out_nodesource is not None else (
            np.empty((self.mesh.nUniqueNodes, 4)) if want_nodesource else None)
        self._check(self.lib.piclas_gpu_deposit(_f(PS), _f(NS)))
        return PS, NS

    def ChargeDensity(self, out=None):
        """PartSource(4,:,:,:) of the last Deposition: all the HDG source term reads (equations/poisson/equation.f90:1043)."""
        rho = out if out is not None else np.empty(self._ps_shape[:-1])
        self._check(self.lib.piclas_gpu_get_charge(_f(rho)))
        return rho

    def PartSourceAsync(self, out):
        """Starts the device -> host copy of PS_N%PartSource of the last Deposition into `out` (page-locked memory) beside the calls
        that follow; `out` is complete after PartSourceWait()."""
        if out.shape != self._ps_shape or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise PiclasGpuError(f"PartSourceAsync: expected a C-contiguous float64 array of shape {self._ps_shape}")
        self._check(self.lib.piclas_gpu_get_partsource_async(_f(out)))

    def PartSourceWait(self):
        self._check(self.lib.piclas_gpu_partsource_wait())

    def KineticEnergy(self):
        """CalcKineticEnergy / CalcNumPartsOfSpec (particle_analyze_tools.f90:709-842) on the device: (Ekin[nSpecies] in J,
        nPart[nSpecies])."""
        ns = len(self.params.ChargeIC)
        E = np.zeros(ns)
        N = np.zeros(ns, dtype=np.int64)
        self._check(self.lib.piclas_gpu_kinetic_energy(_f(E), _l(N)))
        return E, N

    def EmitLattice(self, SpaceIC, iSpec, maxParticleNumber, Amplitude=0.0, WaveNumber=0.0, velocity=(0.0, 0.0, 0.0), append=False) -> int:
        """Initial emission on the device: SpaceIC 'sin_deviation' / 'cos_distribution' (particle_emission_tools.f90:1235-1371) with
        velocityDistribution = constant, localisation by SinglePointToElement (particle_localization.f90:81-190).  Returns the number
        of particles this rank accepted."""
        kinds = {"sin_deviation": 1, "cos_distribution": 2}
        if SpaceIC not in kinds:
            raise PiclasGpuError(f"EmitLattice: SpaceIC={SpaceIC!r}; supported: {sorted(kinds)}")
        n3 = np.ascontiguousarray(maxParticleNumber, dtype=np.int32)
        v3 = np.ascontiguousarray(velocity, dtype=np.float64)
        if n3.shape != (3,) or v3.shape != (3,):
            raise PiclasGpuError("EmitLattice: maxParticleNumber and velocity take three entries")
        ne = C.c_int64(0)
        self._check(self.lib.piclas_gpu_emit_lattice(C.c_int32(kinds[SpaceIC]), C.c_int32(int(iSpec)), n3.ctypes.data_as(C.POINTER(C.c_int32)),
                                                     C.c_double(Amplitude), C.c_double(WaveNumber), _f(v3), C.c_int32(int(bool(append))),
                                                     C.byref(ne)))
        return int(ne.value)

    def SetField(self, E):
        """U_N(iElem)%E(1:3,i,j,k) packed as [nElems,k,j,i,3] after CALL HDG (hdg/elem_mat.f90:709)."""
        E = np.ascontiguousarray(E, dtype=np.float64)
        if E.shape != self._e_shape:
            raise PiclasGpuError(f"SetField: expected shape {self._e_shape}, got {E.shape}")
        self._check(self.lib.piclas_gpu_set_field(_f(E)))

    def PushAndTrack(self, dt: float, iter: int = 0) -> int:
        """timedisc_TimeStepPoissonByBorisLeapfrog.f90:109-215 + :270.  Returns NbrOfLostParticles of this step."""
        nl = C.c_int32(0)
        self._check(self.lib.piclas_gpu_push_track(C.c_double(dt), C.c_int64(iter), C.byref(nl)))
        return nl.value

    def LastTiming(self):
        ms = C.c_double(0.0)
        nl = C.c_int32(0)
        self.lib.piclas_gpu_last_timing(C.byref(ms), C.byref(nl))
        return ms.value, nl.value

    def PhaseTiming(self):
        """ms of [deposit particle kernel, node+DOF kernels, interpolate+push+track kernel, sort+permute]."""
        a = np.zeros(4)
        self.lib.piclas_gpu_phase_timing(_f(a))
        return a

    # ------------------------------------------------------------------------------------------------------
    def TimeStep(self, dt, field_solver, iter=0):
        """One pass of TimeStepPoissonByBorisLeapfrog with the HDG solve left to `field_solver(PartSource) -> E`."""
        PS, _ = self.Deposition(want_nodesource=False)
        E = field_solver(PS)
        self.SetField(E)
        return self.PushAndTrack(dt, iter)
