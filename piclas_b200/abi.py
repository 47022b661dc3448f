"""ctypes mirror of include/piclas_gpu.h (struct layouts only) + marshalling from ParticleMesh.

The structs are the single source of truth for what crosses the C ABI; a Fortran host fills the same
layout through `TYPE, BIND(C)` (piclas_b200/fortran/mod_particle_gpu.f90).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
import numpy as np

from .hostmesh import ParticleMesh, TRIATRACKING, REFMAPPING

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)
c_i64p = C.POINTER(C.c_int64)

DEPO_CVW, DEPO_SF, DEPO_SF_CC, DEPO_SF_ADAPTIVE, DEPO_CVWM = 0, 1, 2, 3, 6
TIMEDISC_LEAPFROG, TIMEDISC_BORIS_LEAPFROG = 509, 508

DEPO_NAMES = {
    "cell_volweight_mean": DEPO_CVWM,
    "shape_function": DEPO_SF,
    "shape_function_cc": DEPO_SF_CC,
    "shape_function_adaptive": DEPO_SF_ADAPTIVE,
}
TRACKING_NAMES = {"refmapping": REFMAPPING, "triatracking": TRIATRACKING}


class pgpu_mesh_t(C.Structure):
    _fields_ = [
        ("nGlobalElems", C.c_int32), ("nSides", C.c_int32), ("nNonUniqueNodes", C.c_int32),
        ("nUniqueGlobalNodes", C.c_int32),
        ("NGeo", C.c_int32), ("N", C.c_int32), ("offsetElem", C.c_int32), ("nElems", C.c_int32),
        ("elemInfoSize", C.c_int32), ("sideInfoSize", C.c_int32),
        ("ElemInfo", c_i32p), ("SideInfo", c_i32p), ("NodeCoords", c_f64p), ("NodeInfo", c_i32p),
        ("ElemNodeID", c_i32p), ("ElemSideNodeID", c_i32p), ("ConcaveElemSide", c_i32p),
        ("XCL_NGeo", c_f64p), ("dXCL_NGeo", c_f64p), ("XiCL_NGeo", c_f64p), ("wBaryCL_NGeo", c_f64p),
        ("ElemBaryNGeo", c_f64p), ("ElemRadius2NGeo", c_f64p), ("XiEtaZetaBasis", c_f64p),
        ("slenXiEtaZetaBasis", c_f64p),
        ("xGP", c_f64p), ("wGP", c_f64p), ("wBary", c_f64p),
        ("Elem_xGP", c_f64p), ("ElemsJ", c_f64p),
        ("nBCs", C.c_int32), ("bc_kind", c_i32p), ("bc_alpha", c_i32p),
        ("nPeriodicVectors", C.c_int32), ("PeriodicVectors", c_f64p),
        ("Periodic_nNodes", c_i32p), ("Periodic_offsetNode", c_i32p), ("Periodic_Nodes", c_i32p),
        ("nPeriodicNodesTotal", C.c_int32), ("NodeVolume", c_f64p),
        ("FIBGMdeltas", C.c_double * 3), ("xyzminglob", C.c_double * 3), ("xyzmaxglob", C.c_double * 3),
        ("FIBGMmin", C.c_int32 * 3), ("FIBGMmax", C.c_int32 * 3),
        ("FIBGM_nElems", c_i32p), ("FIBGM_offsetElem", c_i32p), ("FIBGM_Element", c_i32p),
        ("nFIBGMElemsTotal", C.c_int32),
        ("ElemEpsOneCell", c_f64p), ("ElemToBCSides", c_i32p), ("SideBCMetrics", c_f64p),
        ("nBCSidesTotal", C.c_int32),
        ("SideType", c_i32p), ("SideNormVec", c_f64p), ("SideDistance", c_f64p),
        ("BaseVectors0", c_f64p), ("BaseVectors1", c_f64p), ("BaseVectors2", c_f64p),
        ("BaseVectorsScale", c_f64p),
        ("SFElemr2", c_f64p), ("ElemRadiusNGeo", c_f64p), ("ElemToBGM", c_i32p), ("BaseVectors3", c_f64p),
    ]


class pgpu_params_t(C.Structure):
    _fields_ = [
        ("TrackingMethod", C.c_int32), ("RefMappingGuess", C.c_int32), ("RefMappingEps", C.c_double),
        ("CartesianPeriodic", C.c_int32), ("TimeDiscMethod", C.c_int32),
        ("DoInterpolation", C.c_int32), ("DoDeposition", C.c_int32), ("DepositionType", C.c_int32),
        ("externalField", C.c_double * 6), ("c2_inv", C.c_double),
        ("nSpecies", C.c_int32), ("ChargeIC", c_f64p), ("MassIC", c_f64p), ("MacroParticleFactor", c_f64p),
        ("r_sf", C.c_double), ("alpha_sf", C.c_int32), ("dim_sf", C.c_int32), ("dim_sf_dir", C.c_int32),
        ("sfDepo3D", C.c_int32), ("w_sf", C.c_double), ("dimFactorSF", C.c_double),
        ("device", C.c_int32), ("myRank", C.c_int32), ("nRanks", C.c_int32),
        ("maxParticleNumber", C.c_int64), ("carryParticleIDs", C.c_int32), ("arithmetic", C.c_int32),
        ("PartLorentzType", C.c_int32), ("NoDirichletDeposition", C.c_int32), ("DoDielectricSurfaceCharge", C.c_int32),
    ]


# speed of light as the reference (globals_vars.f90:87-90): c = 299792458, c2_inv = 1/c^2
C0 = 299792458.0
C2_INV = 1.0 / (C0 * C0)


@dataclass
class Params:
    """Already-parsed parameter.ini values that matter on the particle path (pgpu_params_t)."""
    TrackingMethod: int = TRIATRACKING
    RefMappingGuess: int = 1
    RefMappingEps: float = 1e-4
    CartesianPeriodic: int = 0
    TimeDiscMethod: int = TIMEDISC_BORIS_LEAPFROG
    DoInterpolation: int = 1
    DoDeposition: int = 1
    DepositionType: int = DEPO_CVWM
    externalField: tuple = (0.0,) * 6
    c2_inv: float = C2_INV
    ChargeIC: tuple = (-1.60217653e-19,)
    MassIC: tuple = (9.1093826e-31,)
    MacroParticleFactor: tuple = (1.0,)
    r_sf: float = 0.0
    alpha_sf: int = 2
    dim_sf: int = 3
    dim_sf_dir: int = 1
    sfDepo3D: int = 1
    w_sf: float = 0.0
    dimFactorSF: float = 1.0
    device: int = 0
    myRank: int = 0
    nRanks: int = 1
    maxParticleNumber: int = 0
    carryParticleIDs: int = 0
    arithmetic: int = 0
    PartLorentzType: int = 0            # only 0 (non-relativistic) is implemented; anything else fails at init
    NoDirichletDeposition: int = 0      # .NOT. PIC-DoDirichletDeposition
    DoDielectricSurfaceCharge: int = 0


def _p(arr, typ):
    if arr is None:
        return C.cast(None, typ)
    return arr.ctypes.data_as(typ)


class Marshalled:
    """Owns contiguous numpy copies for the lifetime of the C structs that point into them."""

    def __init__(self, mesh: ParticleMesh, params: Params, offsetElem: int = 0, nElems: int | None = None):
        self.keep = []
        m = pgpu_mesh_t()
        m.nGlobalElems = mesh.nElems
        m.nSides = mesh.nSides
        m.nNonUniqueNodes = mesh.nNonUniqueNodes
        m.nUniqueGlobalNodes = mesh.nUniqueNodes
        m.NGeo = mesh.NGeo
        m.N = mesh.N
        m.offsetElem = offsetElem
        m.nElems = mesh.nElems if nElems is None else nElems
        m.elemInfoSize = mesh.ElemInfo.shape[1]
        m.sideInfoSize = mesh.SideInfo.shape[1]

        def i32(a):
            b = np.ascontiguousarray(a, dtype=np.int32)
            self.keep.append(b)
            return _p(b, c_i32p)

        def f64(a):
            b = np.ascontiguousarray(a, dtype=np.float64)
            self.keep.append(b)
            return _p(b, c_f64p)

        m.ElemInfo = i32(mesh.ElemInfo)
        m.SideInfo = i32(mesh.SideInfo)
        m.NodeCoords = f64(mesh.NodeCoords)
        m.NodeInfo = i32(mesh.NodeInfo)
        m.ElemNodeID = i32(mesh.ElemNodeID)
        m.ElemSideNodeID = i32(mesh.ElemSideNodeID)
        m.ConcaveElemSide = i32(mesh.ConcaveElemSide)
        m.XCL_NGeo = f64(mesh.XCL_NGeo)
        m.dXCL_NGeo = f64(mesh.dXCL_NGeo)
        m.XiCL_NGeo = f64(mesh.XiCL_NGeo)
        m.wBaryCL_NGeo = f64(mesh.wBaryCL_NGeo)
        m.ElemBaryNGeo = f64(mesh.ElemBaryNGeo)
        m.ElemRadius2NGeo = f64(mesh.ElemRadius2NGeo)
        m.XiEtaZetaBasis = f64(mesh.XiEtaZetaBasis)
        m.slenXiEtaZetaBasis = f64(mesh.slenXiEtaZetaBasis)
        m.xGP = f64(mesh.xGP)
        m.wGP = f64(mesh.wGP)
        m.wBary = f64(mesh.wBary)
        m.Elem_xGP = f64(mesh.Elem_xGP)
        m.ElemsJ = f64(mesh.sJ)
        m.nBCs = mesh.nBCs
        m.bc_kind = i32(mesh.bc_kind)
        m.bc_alpha = i32(mesh.bc_alpha)
        m.nPeriodicVectors = mesh.nPeriodicVectors
        m.PeriodicVectors = f64(mesh.PeriodicVectors if mesh.nPeriodicVectors else np.zeros((1, 3)))
        m.Periodic_nNodes = i32(mesh.Periodic_nNodes)
        m.Periodic_offsetNode = i32(mesh.Periodic_offsetNode)
        m.Periodic_Nodes = i32(mesh.Periodic_Nodes if mesh.Periodic_Nodes.size else np.zeros(1))
        m.nPeriodicNodesTotal = int(mesh.Periodic_Nodes.size)
        m.NodeVolume = f64(mesh.NodeVolume)
        for d in range(3):
            m.xyzminglob[d] = float(mesh.xyz_min[d])
            m.xyzmaxglob[d] = float(mesh.xyz_max[d])
        ex = mesh.extra
        if "FIBGM" in ex:
            fb = ex["FIBGM"]
            for d in range(3):
                m.FIBGMdeltas[d] = float(fb["deltas"][d])
                m.FIBGMmin[d] = int(fb["min"][d])
                m.FIBGMmax[d] = int(fb["max"][d])
            m.FIBGM_nElems = i32(fb["nElems"])
            m.FIBGM_offsetElem = i32(fb["offsetElem"])
            m.FIBGM_Element = i32(fb["Element"])
            m.nFIBGMElemsTotal = int(fb["Element"].size)
        for name, conv in (("ElemEpsOneCell", f64), ("ElemToBCSides", i32), ("SideBCMetrics", f64),
                           ("SideType", i32), ("SideNormVec", f64), ("SideDistance", f64),
                           ("BaseVectors0", f64), ("BaseVectors1", f64), ("BaseVectors2", f64),
                           ("BaseVectorsScale", f64), ("SFElemr2", f64), ("ElemRadiusNGeo", f64), ("ElemToBGM", i32),
                           ("BaseVectors3", f64)):
            if name in ex:
                setattr(m, name, conv(ex[name]))
        if "SideBCMetrics" in ex:
            m.nBCSidesTotal = int(ex["SideBCMetrics"].shape[0])
        self.mesh = m

        p = pgpu_params_t()
        for k in ("TrackingMethod", "RefMappingGuess", "RefMappingEps", "CartesianPeriodic", "TimeDiscMethod",
                  "DoInterpolation", "DoDeposition", "DepositionType", "c2_inv", "r_sf", "alpha_sf", "dim_sf",
                  "dim_sf_dir", "sfDepo3D", "w_sf", "dimFactorSF", "device", "myRank", "nRanks",
                  "maxParticleNumber", "carryParticleIDs", "arithmetic", "PartLorentzType", "NoDirichletDeposition",
                  "DoDielectricSurfaceCharge"):
            setattr(p, k, getattr(params, k))
        for d in range(6):
            p.externalField[d] = float(params.externalField[d])
        p.nSpecies = len(params.ChargeIC)
        p.ChargeIC = f64(np.array(params.ChargeIC))
        p.MassIC = f64(np.array(params.MassIC))
        p.MacroParticleFactor = f64(np.array(params.MacroParticleFactor))
        self.params = p
