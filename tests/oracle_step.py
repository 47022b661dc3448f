"""Oracle-backed stand-in for piclas_b200.particle_step.ParticleStep (TEST INFRASTRUCTURE): same methods, the CPU oracle as the
engine, particles kept sorted by element as on the device.  tests/test_dry_run_gpu_tests.py uses it to run the bodies of the
-m gpu tests on the CPU, so that their own code (arguments, shapes, bookkeeping, tolerances) is exercised where no GPU exists.
It is never used by the product path, the benchmark or a -m gpu test."""
import numpy as np

from oracle_lib import Oracle
from piclas_b200 import hostmesh as hm
from piclas_b200.abi import DEPO_CVWM
import piclas_b200.particle_step as ps


class OracleStep:
    def __init__(self, mesh, params, offsetElem=0, nElems=None):
        self.mesh, self.params = mesh, params
        self.orc = Oracle(mesh, params)
        self.ref = params.TrackingMethod == hm.REFMAPPING
        self.PS = np.zeros((0, 6)); self.spec = np.zeros(0, np.int32); self.elem = np.zeros(0, np.int32)
        self.isnew = np.zeros(0, np.int32); self.ids = np.zeros(0, np.int64); self.xi = np.zeros((0, 3))
        self.E = None; self.src = None
        n1 = mesh.N + 1
        self._e_shape = (mesh.nElems, n1, n1, n1, 3); self._ps_shape = (mesh.nElems, n1, n1, n1, 4)
        self.offsetElem, self.nElems = 0, mesh.nElems
    def __enter__(self): return self
    def __exit__(self, *a): self.orc.close()
    def close(self): self.orc.close()
    def _sort(self):
        o = np.argsort(self.elem, kind='stable')
        for k in ('PS', 'spec', 'elem', 'isnew', 'ids', 'xi'):
            setattr(self, k, np.ascontiguousarray(getattr(self, k)[o]))
    def UploadParticles(self, PartState, PartSpecies, GlobalElemID, ParticleInside=None, IsNewPart=None, PartPosRef=None, ids=None, append=False):
        PS = np.ascontiguousarray(PartState, dtype=np.float64); n = len(PS)
        assert PS.shape == (n, 6) and len(PartSpecies) == n and len(GlobalElemID) == n
        spec = np.asarray(PartSpecies, np.int32); elem = np.asarray(GlobalElemID, np.int32)
        isnew = np.zeros(n, np.int32) if IsNewPart is None else np.asarray(IsNewPart, np.int32)
        idv = np.arange(n, dtype=np.int64) if ids is None else np.asarray(ids, np.int64)
        if self.ref:
            xi = self.orc.position_in_ref_elem(PS[:, :3], elem)[0] if PartPosRef is None else np.asarray(PartPosRef, np.float64)
            assert xi.shape == (n, 3)
        else:
            xi = np.zeros((n, 3))
        keep = np.ones(n, bool) if ParticleInside is None else np.asarray(ParticleInside) != 0
        if not append:
            self.PS, self.spec, self.elem, self.isnew, self.ids, self.xi = PS[keep].copy(), spec[keep].copy(), elem[keep].copy(), isnew[keep].copy(), idv[keep].copy(), xi[keep].copy()
        else:
            raise NotImplementedError
        self._sort()
    def NumParticles(self): return len(self.spec)
    def SetField(self, E):
        E = np.ascontiguousarray(E, dtype=np.float64); assert E.shape == self._e_shape, (E.shape, self._e_shape); self.E = E
    def Deposition(self, want_partsource=True, want_nodesource=True, out_partsource=None, out_nodesource=None):
        assert self.params.DoDeposition
        ins = np.ones(len(self.spec), np.int32)
        src, ns = self.orc.deposit(self.PS, self.spec, self.elem, ins, PartPosRef=self.xi if self.ref else None)
        self.src = src
        if out_partsource is not None: out_partsource[...] = src
        return (src if want_partsource else None), (ns if (want_nodesource and self.params.DepositionType == DEPO_CVWM) else None)
    def ChargeDensity(self, out=None):
        if out is not None:
            out[...] = self.src[..., 3]; return out
        return self.src[..., 3].copy()
    def PartSourceAsync(self, out): out[...] = self.src
    def PartSourceWait(self): pass
    def PushAndTrack(self, dt, iter=0):
        assert self.E is not None or not self.params.DoInterpolation
        n = len(self.spec); ins = np.ones(n, np.int32)
        E = self.E if self.E is not None else np.zeros(self._e_shape)
        nl, _, _ = self.orc.push_track(dt, self.PS, self.spec, self.elem, ins, self.isnew, E, PartPosRef=self.xi if self.ref else None)
        k = ins != 0
        for a in ('PS', 'spec', 'elem', 'isnew', 'ids', 'xi'): setattr(self, a, getattr(self, a)[k])
        self._sort()
        return nl
    def DownloadParticles(self, want_ref=False):
        return dict(PartState=self.PS.copy(), PartSpecies=self.spec.copy(), GlobalElemID=self.elem.copy(), PartPosRef=self.xi.copy() if want_ref else None,
                    ids=self.ids.copy() if self.params.carryParticleIDs else None)
    def KineticEnergy(self):
        p = self.params; ns = len(p.ChargeIC); c2 = 1.0 / p.c2_inv
        v2 = (self.PS[:, 3:] ** 2).sum(axis=1); m = np.asarray(p.MassIC)[self.spec - 1]; mpf = np.asarray(p.MacroParticleFactor)[self.spec - 1]
        ek = np.where(v2 < (1e6 / 299792458.0) ** 2 * c2, 0.5 * m * v2, (1.0 / np.sqrt(1.0 - v2 / c2) - 1.0) * m * c2) * mpf
        return np.array([ek[self.spec == s + 1].sum() for s in range(ns)]), np.array([(self.spec == s + 1).sum() for s in range(ns)], dtype=np.int64)
    FillParticleData = ps.ParticleStep.FillParticleData
