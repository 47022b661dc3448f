"""ctypes binding of oracle/_build/liboracle.so — the CPU checker (test infrastructure only)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

from piclas_b200.abi import Marshalled, Params, pgpu_mesh_t, pgpu_params_t, c_f64p, c_i32p

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def _load(fast=False):
    name = "liboracle_fast.so" if fast else "liboracle.so"
    path = os.path.join(ORACLE_DIR, "_build", name)
    src = os.path.join(ORACLE_DIR, "piclas_oracle.cpp")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        build_oracle()
    lib = C.CDLL(path)
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_create.argtypes = [C.POINTER(pgpu_mesh_t), C.POINTER(pgpu_params_t)]
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_last_error.restype = C.c_char_p
    lib.oracle_last_error.argtypes = [C.c_void_p]
    lib.oracle_deposited_charge.restype = C.c_double
    lib.oracle_deposited_charge.argtypes = [C.c_void_p, c_f64p]
    return lib


def _f(a):
    return a.ctypes.data_as(c_f64p) if a is not None else C.cast(None, c_f64p)


def _i(a):
    return a.ctypes.data_as(c_i32p) if a is not None else C.cast(None, c_i32p)


class Oracle:
    def __init__(self, mesh, params: Params, fast=False, offsetElem=0, nElems=None):
        self.lib = _load(fast)
        self.msh = mesh
        self.params = params
        self.mar = Marshalled(mesh, params, offsetElem=offsetElem, nElems=nElems)
        self.h = C.c_void_p(self.lib.oracle_create(C.byref(self.mar.mesh), C.byref(self.mar.params)))
        self.nloc = mesh.nElems if nElems is None else nElems

    def close(self):
        if self.h:
            self.lib.oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise RuntimeError("oracle: rc=%d %s" % (rc, self.lib.oracle_last_error(self.h).decode()))

    def lagrange(self, x, xgp, wbary):
        L = np.zeros(len(xgp))
        self.lib.oracle_lagrange_polys(C.c_double(x), C.c_int(len(xgp) - 1), _f(np.ascontiguousarray(xgp)),
                                       _f(np.ascontiguousarray(wbary)), _f(L))
        return L

    def position_in_ref_elem(self, x, elem, force=True):
        x = np.ascontiguousarray(x, dtype=np.float64)
        elem = np.ascontiguousarray(elem, dtype=np.int32)
        xi = np.zeros_like(x)
        suc = np.zeros(len(elem), dtype=np.int32)
        bad = self.lib.oracle_position_in_ref_elem(self.h, C.c_int64(len(elem)), _f(x), _i(elem), C.c_int(int(force)),
                                                   _f(xi), _i(suc))
        return xi, suc, bad

    def inside(self, x, elem):
        x = np.ascontiguousarray(x, dtype=np.float64)
        elem = np.ascontiguousarray(elem, dtype=np.int32)
        ins = np.zeros(len(elem), dtype=np.int32)
        det = np.zeros((len(elem), 6, 2))
        self.lib.oracle_inside_quad3d(self.h, C.c_int64(len(elem)), _f(x), _i(elem), _i(ins), _f(det))
        return ins, det

    def locate(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        el = np.zeros(len(x), dtype=np.int32)
        self.lib.oracle_locate(self.h, C.c_int64(len(x)), _f(x), _i(el))
        return el

    def interpolate(self, PartState, elem, E, PartPosRef=None):
        n = len(elem)
        F = np.zeros((n, 6))
        self._check(self.lib.oracle_interpolate(self.h, C.c_int64(n), _f(PartState), _i(elem), _f(PartPosRef), _f(E), _f(F)))
        return F

    def push_track(self, dt, PartState, PartSpecies, GlobalElemID, ParticleInside, IsNewPart, E, PartPosRef=None,
                   threads=0):
        """In-place on the given arrays (as the Fortran module globals). Returns (nLost, LastPartPos, Field)."""
        n = len(PartSpecies)
        last = np.zeros((n, 3))
        F = np.zeros((n, 6))
        nl = C.c_int32(0)
        if threads and threads > 1:
            rc = self.lib.oracle_push_track_mt(self.h, C.c_int(threads), C.c_double(dt), C.c_int64(n), _f(PartState),
                                               _f(last), _i(PartSpecies), _i(GlobalElemID), _i(ParticleInside),
                                               _i(IsNewPart), _f(PartPosRef), _f(E), C.byref(nl))
        else:
            rc = self.lib.oracle_push_track(self.h, C.c_double(dt), C.c_int64(n), _f(PartState), _f(last),
                                            _i(PartSpecies), _i(GlobalElemID), _i(ParticleInside), _i(IsNewPart),
                                            _f(PartPosRef), _f(E), _f(F), C.byref(nl))
        self._check(rc)
        return nl.value, last, F

    def deposit(self, PartState, PartSpecies, GlobalElemID, ParticleInside, PartPosRef=None, threads=0):
        n = len(PartSpecies)
        n1 = self.msh.N + 1
        PS = np.zeros((self.nloc, n1, n1, n1, 4))
        NS = np.zeros((self.msh.nUniqueNodes, 4))
        if threads and threads > 1:
            rc = self.lib.oracle_deposit_mt(self.h, C.c_int(threads), C.c_int64(n), _f(PartState), _i(PartSpecies),
                                            _i(GlobalElemID), _i(ParticleInside), _f(PS), _f(NS))
        else:
            rc = self.lib.oracle_deposit(self.h, C.c_int64(n), _f(PartState), _i(PartSpecies), _i(GlobalElemID),
                                         _i(ParticleInside), _f(PartPosRef), _f(PS), _f(NS))
        self._check(rc)
        return PS, NS

    def deposit_raw(self, PartState, PartSpecies, GlobalElemID, ParticleInside):
        NS = np.zeros((self.msh.nUniqueNodes, 4))
        self._check(self.lib.oracle_deposit_raw(self.h, C.c_int64(len(PartSpecies)), _f(PartState), _i(PartSpecies),
                                                _i(GlobalElemID), _i(ParticleInside), _f(NS)))
        return NS

    def deposit_finish(self, NS):
        n1 = self.msh.N + 1
        PS = np.zeros((self.nloc, n1, n1, n1, 4))
        self._check(self.lib.oracle_deposit_finish(self.h, _f(NS), _f(PS)))
        return PS, NS

    def deposited_charge(self, PartSource):
        return float(self.lib.oracle_deposited_charge(self.h, _f(PartSource)))
