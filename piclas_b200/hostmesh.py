"""Host-side particle-mesh tables for straight-sided (NGeo=1) hexahedral meshes.

In a PICLas run these tables are produced once by the Fortran host
(`InitParticleMesh`, reference src/particles/particle_mesh/particle_mesh.f90:141-531, and
`InitializeDeposition`, src/particles/pic/deposition/pic_depo.f90:83-626) and handed to the
device layer through `piclas_gpu_init` (include/piclas_gpu.h).  This module builds the same
tables for stand-alone use (tests, bench, oracle) from a conforming hex mesh given as corner
coordinates + CGNS-ordered element connectivity, which covers every mesh of the five configs
(`hopr.ini` Corner/nElems boxes).  Memory layout: every array is C-contiguous with the Fortran
index order reversed, i.e. byte-identical to the Fortran column-major array it mirrors.

Builders restated here (the construction itself stays a host job):
  ElemInfo/SideInfo incl. SIDE_ELEMID/LOCALID/NBSIDEID  particle_mesh_readin.f90:386-419, piclas.h:149-176
  ElemNodeID, CGNS corner/side node maps                 mesh/mesh_tools.f90:634-769
  ElemSideNodeID, ConcaveElemSide                        particle_mesh_tools.f90:1800-1954
  XCL_NGeo, dXCL_NGeo, XiCL/wBaryCL                      mesh/metrics.f90:250-336
  ElemBaryNGeo, ElemRadius2NGeo (TriaTracking)           particle_mesh_build.f90:217-302
  XiEtaZetaBasis, slenXiEtaZetaBasis                     particle_mesh_build.f90:407-433
  NodeVolume                                             pic_depo_tools.f90:224-355
  periodic node partners (result of)                     pic_depo.f90:1218-2089
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np
from scipy.spatial import cKDTree

from . import basis

# piclas.h:204-209
ZETA_MINUS, ETA_MINUS, XI_PLUS, ETA_PLUS, XI_MINUS, ZETA_PLUS = 1, 2, 3, 4, 5, 6
# tracking ids piclas.h:357-359
REFMAPPING, TRACING, TRIATRACKING = 1, 2, 3
# PartBound%TargetBoundCond values used on this path (particle_boundary_condition.f90:167-214)
BC_OPEN, BC_REFLECTIVE, BC_PERIODIC = 1, 2, 3

# CGNS corner c (0-based) -> tensor (i,j,k) of the NGeo=1 node block (mesh_tools.f90:742-749)
_CGNS_IJK = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0],
                      [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], dtype=np.int64)
# tensor-local node number (0-based, i fastest) of CGNS corner c: CNS(:)-1
_CNS0 = _CGNS_IJK[:, 0] + 2 * _CGNS_IJK[:, 1] + 4 * _CGNS_IJK[:, 2]
# NodeMapCGNS(1:4, locSide) expressed in 0-based CGNS corner numbers (mesh_tools.f90:761-766)
_NODEMAP_CGNS0 = np.array([[0, 3, 2, 1],
                           [0, 1, 5, 4],
                           [1, 2, 6, 5],
                           [2, 3, 7, 6],
                           [0, 4, 7, 3],
                           [4, 5, 6, 7]], dtype=np.int64)


@dataclass
class ParticleMesh:
    """All device-layer inputs that describe mesh + basis (see include/piclas_gpu.h: pgpu_mesh_t)."""
    N: int
    NGeo: int
    tracking: int
    nElems: int
    nSides: int
    nNonUniqueNodes: int
    nUniqueNodes: int
    ElemInfo: np.ndarray          # (nElems, 8) int32
    SideInfo: np.ndarray          # (nSides, 8) int32
    NodeCoords: np.ndarray        # (nNonUniqueNodes, 3) f64
    NodeInfo: np.ndarray          # (nNonUniqueNodes,) int32, 1-based unique node id
    ElemNodeID: np.ndarray        # (nElems, 8) int32, 1-based non-unique node index, CGNS order
    ElemSideNodeID: np.ndarray    # (nElems, 6, 4) int32, 0-based non-unique node index
    ConcaveElemSide: np.ndarray   # (nElems, 6) int32 (Fortran LOGICAL)
    XCL_NGeo: np.ndarray          # (nElems, NGeo+1, NGeo+1, NGeo+1, 3)
    dXCL_NGeo: np.ndarray         # (nElems, NGeo+1, NGeo+1, NGeo+1, 3, 3)  [e,k,j,i,nn,dd]
    XiCL_NGeo: np.ndarray
    wBaryCL_NGeo: np.ndarray
    ElemBaryNGeo: np.ndarray      # (nElems, 3)
    ElemRadius2NGeo: np.ndarray   # (nElems,)
    XiEtaZetaBasis: np.ndarray    # (nElems, 6, 3)
    slenXiEtaZetaBasis: np.ndarray  # (nElems, 6)
    xGP: np.ndarray
    wGP: np.ndarray
    wBary: np.ndarray
    Elem_xGP: np.ndarray          # (nElems, N+1, N+1, N+1, 3)  [e,k,j,i,:]
    sJ: np.ndarray                # (nElems, N+1, N+1, N+1)
    nBCs: int
    bc_kind: np.ndarray           # (nBCs,) int32  PartBound%TargetBoundCond(MapToPartBC(BCID))
    bc_alpha: np.ndarray          # (nBCs,) int32  BoundaryType(BCID,BC_ALPHA)
    PeriodicVectors: np.ndarray   # (nPV, 3)
    Periodic_nNodes: np.ndarray   # (nUniqueNodes,) int32
    Periodic_offsetNode: np.ndarray  # (nUniqueNodes,) int32
    Periodic_Nodes: np.ndarray    # (sum,) int32 1-based unique ids
    NodeVolume: np.ndarray        # (nUniqueNodes,)
    xyz_min: np.ndarray
    xyz_max: np.ndarray
    unique_coords: np.ndarray     # (nUniqueNodes, 3) convenience (not part of the reference tables)
    elem_nodes: np.ndarray        # (nElems, 8) 0-based unique ids, CGNS order (convenience)
    extra: dict = field(default_factory=dict)

    @property
    def nPeriodicVectors(self) -> int:
        return int(self.PeriodicVectors.shape[0])


def _det3(a):
    """getDet (eval_xyz.f90:448-468) on arrays a[...,r,c]."""
    return ((a[..., 0, 0] * a[..., 1, 1] - a[..., 0, 1] * a[..., 1, 0]) * a[..., 2, 2]
            + (a[..., 0, 1] * a[..., 1, 2] - a[..., 0, 2] * a[..., 1, 1]) * a[..., 2, 0]
            + (a[..., 0, 2] * a[..., 1, 0] - a[..., 0, 0] * a[..., 1, 2]) * a[..., 2, 1])


def build_mesh(coords, elem_nodes, N, bcs, bc_of_face, periodic_vectors=(),
               tracking=TRIATRACKING, hopr_sides=None) -> ParticleMesh:
    """Build all particle-mesh tables.

    hopr_sides    optional (6*nE, 5) SideInfo of a HOPR mesh file (SideType, SIDE_ID, nbElemID, 10*nbLocSide+flip, BCID): side
                  ids, master/slave, flips and BC ids are then taken from the file (they fix which diagonal splits a shared
                  quadrilateral and which element owns a planar face, particle_mesh_readin.f90:386-419) instead of being derived
                  from the geometry; bc_of_face is not called.

    coords        (nU,3) unique node coordinates
    elem_nodes    (nE,8) 0-based unique node ids per element in CGNS corner order
    bcs           list of (kind, alpha): kind in {BC_OPEN, BC_REFLECTIVE, BC_PERIODIC},
                  alpha = signed 1-based periodic-vector id (BoundaryType(:,BC_ALPHA)), 0 otherwise
    bc_of_face    callable(centroids (m,3), face_node_ids (m,4)) -> (m,) 1-based BCID of every unmatched face
    """
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    elem_nodes = np.ascontiguousarray(elem_nodes, dtype=np.int64)
    nE = elem_nodes.shape[0]
    nU = coords.shape[0]
    nS = 6 * nE
    PV = np.ascontiguousarray(np.asarray(periodic_vectors, dtype=np.float64).reshape(-1, 3))
    bc_kind = np.array([b[0] for b in bcs], dtype=np.int32)
    bc_alpha = np.array([b[1] for b in bcs], dtype=np.int32)

    # ---- per-element node block in tensor order (HOPR NodeCoords of an NGeo=1 hex) -------------------
    Xc = coords[elem_nodes]                                   # (nE, 8 cgns, 3)
    NodeCoords = np.empty((nE, 8, 3))
    NodeCoords[:, _CNS0, :] = Xc
    NodeInfo = np.empty((nE, 8), dtype=np.int32)
    NodeInfo[:, _CNS0] = (elem_nodes + 1).astype(np.int32)
    first_node = 8 * np.arange(nE, dtype=np.int64)
    ElemNodeID = (first_node[:, None] + (_CNS0[None, :] + 1)).astype(np.int32)

    ElemInfo = np.zeros((nE, 8), dtype=np.int32)
    ElemInfo[:, 0] = 108
    ElemInfo[:, 1] = 1
    ElemInfo[:, 2] = 6 * np.arange(nE)
    ElemInfo[:, 3] = 6 * (np.arange(nE) + 1)
    ElemInfo[:, 4] = first_node
    ElemInfo[:, 5] = first_node + 8
    ElemInfo[:, 6] = 0      # ELEM_RANK (filled by partition())
    ElemInfo[:, 7] = 1      # ELEM_HALOFLAG

    # ---- face connectivity ---------------------------------------------------------------------------
    face_nodes = elem_nodes[:, _NODEMAP_CGNS0]                # (nE, 6, 4) unique ids in NodeMap order
    fn = face_nodes.reshape(nS, 4)
    key = np.sort(fn, axis=1)
    order = np.lexsort((key[:, 3], key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    same = np.all(ks[1:] == ks[:-1], axis=1)
    partner = np.full(nS, -1, dtype=np.int64)
    ia = order[:-1][same]
    ib = order[1:][same]
    partner[ia] = ib
    partner[ib] = ia
    if np.any(same[1:] & same[:-1]):
        raise ValueError("non-manifold mesh: a face is shared by more than two elements")

    centroid = coords[fn].mean(axis=1)
    shift = np.zeros((nS, 3))
    sidx = np.arange(nS)
    if hopr_sides is not None:
        H = np.asarray(hopr_sides, dtype=np.int64)
        if H.shape != (nS, 5):
            raise ValueError("hopr_sides must have shape (6*nElems, 5)")
        if np.any(H[:, 0] > 100) or np.any(H[:, 2] < 0):
            raise ValueError("mortar sides are not supported")
        bcid = H[:, 4].astype(np.int32)
        side_uid = H[:, 1].copy()
        flip = np.where(side_uid > 0, 0, H[:, 3] % 10)
        has_nb = H[:, 2] > 0
        nb_loc = np.where(has_nb, H[:, 3] // 10, 0)
        partner = np.where(has_nb, 6 * (H[:, 2] - 1) + (nb_loc - 1), -1)
        if np.any(bcid > len(bcs)) or np.any(bcid < 0):
            raise ValueError("BCID outside the boundary list")
        isper = (bcid > 0) & (bc_kind[np.maximum(bcid, 1) - 1] == BC_PERIODIC)
        per = np.nonzero(isper)[0]
        if per.size:
            a = bc_alpha[bcid[per] - 1]
            shift[per] = np.sign(a)[:, None] * PV[np.abs(a) - 1]
            ext = np.linalg.norm(coords.max(axis=0) - coords.min(axis=0))
            if np.any(np.linalg.norm(centroid[per] + shift[per] - centroid[partner[per]], axis=1) > 1e-8 * ext):
                raise ValueError("periodic vectors do not map the periodic sides onto their partners")
    else:
        bcid = np.zeros(nS, dtype=np.int32)
        bnd = np.nonzero(partner < 0)[0]
        if bnd.size:
            bcid[bnd] = np.asarray(bc_of_face(centroid[bnd], fn[bnd]), dtype=np.int32)
            if np.any(bcid[bnd] < 1) or np.any(bcid[bnd] > len(bcs)):
                raise ValueError("bc_of_face returned an invalid BCID")
        # periodic pairing of boundary faces by shifted centroid
        per = bnd[bc_kind[bcid[bnd] - 1] == BC_PERIODIC] if bnd.size else bnd
        if per.size:
            a = bc_alpha[bcid[per] - 1]
            shift[per] = np.sign(a)[:, None] * PV[np.abs(a) - 1]
            tree = cKDTree(centroid[per])
            ext = np.linalg.norm(coords.max(axis=0) - coords.min(axis=0))
            d, j = tree.query(centroid[per] + shift[per])
            if np.any(d > 1e-8 * ext):
                raise ValueError("periodic face without partner")
            tgt = per[j]
            if np.any(bc_alpha[bcid[tgt] - 1] != -a):
                raise ValueError("periodic partner has inconsistent BC_ALPHA")
            partner[per] = tgt

        # master / slave, unique side ids, flip
        has_nb = partner >= 0
        is_master = (~has_nb) | (sidx <= partner)
        side_uid = np.zeros(nS, dtype=np.int64)
        masters = np.nonzero(is_master)[0]
        side_uid[masters] = np.arange(1, masters.size + 1)
        slaves = np.nonzero(~is_master)[0]
        side_uid[slaves] = -side_uid[partner[slaves]]
        flip = np.zeros(nS, dtype=np.int64)
        if slaves.size:
            m = partner[slaves]
            p_first = coords[fn[m, 0]] + shift[m]                 # master's first node seen from the slave side
            ps = coords[fn[slaves]]                               # (ns, 4, 3)
            dist = np.linalg.norm(ps - p_first[:, None, :], axis=2)
            flip[slaves] = np.argmin(dist, axis=1) + 1
        nb_loc = np.where(has_nb, (partner % 6) + 1, 0)
    loc_side = (sidx % 6) + 1

    SideInfo = np.zeros((nS, 8), dtype=np.int32)
    SideInfo[:, 0] = 4                                         # SIDE_TYPE (<=100: not a mortar)
    SideInfo[:, 1] = side_uid                                  # SIDE_ID
    SideInfo[:, 2] = np.where(has_nb, partner // 6 + 1, 0)     # SIDE_NBELEMID
    SideInfo[:, 3] = 10 * nb_loc + flip                        # SIDE_FLIP
    SideInfo[:, 4] = bcid                                      # SIDE_BCID
    SideInfo[:, 5] = sidx // 6 + 1                             # SIDE_ELEMID
    SideInfo[:, 6] = loc_side                                  # SIDE_LOCALID
    # SIDE_NBSIDEID: first side of the neighbour element with the same |SIDE_ID| (readin.f90:399-411)
    nbside = np.zeros(nS, dtype=np.int64)
    if has_nb.any():
        h = np.nonzero(has_nb)[0]
        nbe = partner[h] // 6
        cand = np.abs(side_uid.reshape(nE, 6)[nbe])           # (nh, 6)
        hit = cand == np.abs(side_uid[h])[:, None]
        nbside[h] = 6 * nbe + np.argmax(hit, axis=1) + 1       # 1-based SideInfo index
    SideInfo[:, 7] = nbside

    # ---- ElemSideNodeID / ConcaveElemSide --------------------------------------------------------------
    nstart = np.where(side_uid > 0, 0, np.maximum(0, (SideInfo[:, 3] % 10) - 1)).reshape(nE, 6)
    rot = (nstart[:, :, None] + np.arange(4)[None, None, :]) % 4            # (nE,6,4)
    nm_tensor0 = _CNS0[_NODEMAP_CGNS0]                                       # NodeMap(:,side)-1 (tensor local)
    loc = nm_tensor0[np.arange(6)[None, :, None], rot]                       # (nE,6,4)
    ElemSideNodeID = (first_node[:, None, None] + loc).astype(np.int32)      # 0-based non-unique index
    NC = NodeCoords.reshape(nE * 8, 3)
    P = NC[ElemSideNodeID]                                                   # (nE,6,4,3)
    A = P[:, :, 0:3, :] - P[:, :, 3:4, :]                                    # A(:,NodeNum) = node - node4
    # detcon (particle_mesh_tools.f90:1918-1921); A(x,n) -> A[..., n-1, x-1]
    detcon = ((A[..., 0, 1] * A[..., 1, 2] - A[..., 0, 2] * A[..., 1, 1]) * A[..., 2, 0]
              + (A[..., 0, 2] * A[..., 1, 0] - A[..., 0, 0] * A[..., 1, 2]) * A[..., 2, 1]
              + (A[..., 0, 0] * A[..., 1, 1] - A[..., 0, 1] * A[..., 1, 0]) * A[..., 2, 2])
    gsid = (sidx + 1).reshape(nE, 6)
    Concave = (detcon < 0) | ((detcon == 0.0) & (gsid < SideInfo[:, 7].reshape(nE, 6)))
    ConcaveElemSide = Concave.astype(np.int32)

    # ---- geometry for the Newton reference mapping ------------------------------------------------------
    NGeo = 1
    XiCL = basis.cheb_gauss_lobatto_nodes(NGeo)
    wBaryCL = basis.barycentric_weights(XiCL)
    XCL = np.ascontiguousarray(NodeCoords.reshape(nE, 2, 2, 2, 3))           # [e,k,j,i,:]
    D = basis.poly_derivative_matrix(XiCL)
    dXCL = np.zeros((nE, 2, 2, 2, 3, 3))                                     # [e,k,j,i,nn,dd]
    for k in range(2):
        for j in range(2):
            for i in range(2):
                for ll in range(2):
                    dXCL[:, k, j, i, :, 0] += D[i, ll] * XCL[:, k, j, ll, :]
                    dXCL[:, k, j, i, :, 1] += D[j, ll] * XCL[:, k, ll, i, :]
                    dXCL[:, k, j, i, :, 2] += D[k, ll] * XCL[:, ll, j, i, :]

    if tracking == TRIATRACKING:
        # BuildElementRadiusTria (particle_mesh_build.f90:286-301): mean of the first 8 element nodes
        xs = np.zeros((nE, 3))
        for n in range(8):
            xs = xs + NodeCoords[:, n, :]
        Bary = xs / 8.0
        rad = np.zeros(nE)
        for n in range(8):
            v = NodeCoords[:, n, :] - Bary
            rad = np.maximum(rad, np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]))
        Radius2 = rad * rad
    else:
        # BuildElementOriginShared (particle_mesh_build.f90:1387-1470): X(xi=0)
        L0 = basis.lagrange_polys(0.0, XiCL, wBaryCL)
        Bary = np.zeros((nE, 3))
        for k in range(2):
            for j in range(2):
                for i in range(2):
                    Bary = Bary + XCL[:, k, j, i, :] * L0[i] * L0[j] * L0[k]
        rad = np.zeros(nE)
        for k in range(2):
            for j in range(2):
                for i in range(2):
                    v = XCL[:, k, j, i, :] - Bary
                    rad = np.maximum(rad, np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]))
        Radius2 = rad * rad

    XiDirs = np.array([[1., 0., 0.], [0., 1., 0.], [0., 0., 1.],
                       [-1., 0., 0.], [0., -1., 0.], [0., 0., -1.]])
    XEZ = np.zeros((nE, 6, 3))
    slen = np.zeros((nE, 6))
    for d in range(6):
        Lg = [basis.lagrange_polys(XiDirs[d, c], XiCL, wBaryCL) for c in range(3)]
        xPos = np.zeros((nE, 3))
        for k in range(2):
            for j in range(2):
                for i in range(2):
                    xPos = xPos + XCL[:, k, j, i, :] * Lg[0][i] * Lg[1][j] * Lg[2][k]
        XEZ[:, d, :] = xPos - Bary
        slen[:, d] = 1.0 / (XEZ[:, d, 0] * XEZ[:, d, 0] + XEZ[:, d, 1] * XEZ[:, d, 1] + XEZ[:, d, 2] * XEZ[:, d, 2])

    # ---- solution basis, Gauss-point coordinates, inverse Jacobian ----------------------------------------
    xGP, wGP = basis.legendre_gauss_nodes_weights(N)
    wBary = basis.barycentric_weights(xGP)
    LG = np.array([basis.lagrange_polys(x, XiCL, wBaryCL) for x in xGP])    # (N+1, 2)
    # trilinear blend weights of the 8 CL corners at every Gauss point: W[(k,j,i), (kc,jc,ic)]
    W = (LG[:, None, None, :, None, None] * LG[None, :, None, None, :, None]
         * LG[None, None, :, None, None, :]).reshape((N + 1) ** 3, 8)
    Elem_xGP = np.einsum("qc,ecx->eqx", W, XCL.reshape(nE, 8, 3), optimize=True).reshape(nE, N + 1, N + 1, N + 1, 3)
    Jac = np.einsum("qc,ecx->eqx", W, dXCL.reshape(nE, 8, 9), optimize=True).reshape(nE, N + 1, N + 1, N + 1, 3, 3)
    detJ = _det3(Jac)
    # metrics.f90:255-372: the reference evaluates DetJac at the Gauss points of degree NGeoRef = 3 NGeo and brings it to the solution
    # basis with a MODAL Vandermonde matrix (GetVandermonde(..., modal=.TRUE.), interpolation.f90:395-403): for N >= NGeoRef that is the
    # value at the Gauss point (above); for N < NGeoRef it is the L2 projection onto degree N, which differs on non-parallelepiped
    # elements.  sJ enters NodeVolume (pic_depo_tools.f90:282-338) and CalcDepositedCharge.
    NGeoRef = 3
    if N < NGeoRef:
        xr, _ = basis.legendre_gauss_nodes_weights(NGeoRef)
        LGr = np.array([basis.lagrange_polys(x, XiCL, wBaryCL) for x in xr])            # (4, 2)
        Wr = (LGr[:, None, None, :, None, None] * LGr[None, :, None, None, :, None]
              * LGr[None, None, :, None, None, :]).reshape((NGeoRef + 1) ** 3, 8)
        Jr = np.einsum("qc,ecx->eqx", Wr, dXCL.reshape(nE, 8, 9), optimize=True).reshape(nE, NGeoRef + 1, NGeoRef + 1, NGeoRef + 1, 3, 3)
        detr = _det3(Jr)                                                                  # [e, k, j, i]
        Vin = np.polynomial.legendre.legvander(xr, NGeoRef)                               # (4, 4) Legendre modes at the input nodes
        Vout = np.polynomial.legendre.legvander(xGP, N)                                   # (N+1, N+1)
        P = Vout @ np.linalg.inv(Vin)[:N + 1, :]                                          # Vdm_Leg_Out(0:N,0:N) sVdm_Leg_In(0:N,0:NGeoRef)
        detJ = np.einsum("ck,bj,ai,ekji->ecba", P, P, P, detr, optimize=True)
    if np.any(detJ <= 0):
        raise ValueError("mesh has elements with non-positive Jacobian")
    sJ = 1.0 / detJ

    # ---- periodic node partners (CSR, 1-based) ---------------------------------------------------------------
    Periodic_nNodes = np.zeros(nU, dtype=np.int32)
    Periodic_offset = np.zeros(nU, dtype=np.int32)
    Periodic_Nodes = np.zeros(0, dtype=np.int32)
    nPV = PV.shape[0]
    if nPV > 0:
        cand = np.unique(fn[per]) if per.size else np.zeros(0, dtype=np.int64)   # only BC nodes have images
        tree = cKDTree(coords[cand]) if cand.size else None
        ext = np.linalg.norm(coords.max(axis=0) - coords.min(axis=0))
        rows, cols = [np.zeros(0, dtype=np.int64)], [np.zeros(0, dtype=np.int64)]
        import itertools
        for s in itertools.product((-1, 0, 1), repeat=nPV):
            if not any(s):
                continue
            sh = np.zeros(3)
            for c, sv in zip(s, PV):
                sh = sh + c * sv
            if tree is None:
                break
            d, j = tree.query(coords[cand] + sh)
            ok = d <= 1e-8 * ext
            rows.append(cand[ok])
            cols.append(cand[j[ok]])
        rows = np.concatenate(rows)
        cols = np.concatenate(cols)
        o = np.lexsort((cols, rows))
        rows, cols = rows[o], cols[o]
        Periodic_nNodes = np.bincount(rows, minlength=nU).astype(np.int32)
        Periodic_offset = (np.cumsum(Periodic_nNodes) - Periodic_nNodes).astype(np.int32)
        Periodic_Nodes = (cols + 1).astype(np.int32)

    # ---- NodeVolume (pic_depo_tools.f90:282-338) -----------------------------------------------------------------
    NodeVolume = np.zeros(nU)
    uid = NodeInfo[np.arange(nE)[:, None], _CNS0[None, :]].astype(np.int64) - 1     # (nE, 8 cgns)
    sgn = 2.0 * _CGNS_IJK - 1.0                                                       # (8,3) +-1
    F = np.zeros((N + 1, N + 1, N + 1, 8))                                            # [k,j,i,c]
    for k in range(N + 1):
        for j in range(N + 1):
            for i in range(N + 1):
                for c in range(8):
                    F[k, j, i, c] = ((1. + sgn[c, 0] * xGP[i]) * (1. + sgn[c, 1] * xGP[j]) * (1. + sgn[c, 2] * xGP[k])
                                     * wGP[i] * wGP[j] * wGP[k] / 8.)
    vol_ec = (1.0 / sJ).reshape(nE, -1) @ F.reshape(-1, 8)                             # (nE, 8)
    for c in range(8):
        NodeVolume += np.bincount(uid[:, c], weights=vol_ec[:, c], minlength=nU)
    if nPV > 0 and Periodic_Nodes.size:
        add = np.zeros(nU)
        src = np.repeat(np.arange(nU), Periodic_nNodes)
        add = np.bincount(src, weights=NodeVolume[Periodic_Nodes - 1], minlength=nU)
        NodeVolume = NodeVolume + add

    return ParticleMesh(
        N=N, NGeo=NGeo, tracking=tracking, nElems=nE, nSides=nS, nNonUniqueNodes=8 * nE, nUniqueNodes=nU,
        ElemInfo=ElemInfo, SideInfo=SideInfo, NodeCoords=np.ascontiguousarray(NC),
        NodeInfo=np.ascontiguousarray(NodeInfo.reshape(-1)), ElemNodeID=np.ascontiguousarray(ElemNodeID),
        ElemSideNodeID=np.ascontiguousarray(ElemSideNodeID), ConcaveElemSide=np.ascontiguousarray(ConcaveElemSide),
        XCL_NGeo=XCL, dXCL_NGeo=np.ascontiguousarray(dXCL), XiCL_NGeo=XiCL, wBaryCL_NGeo=wBaryCL,
        ElemBaryNGeo=np.ascontiguousarray(Bary), ElemRadius2NGeo=np.ascontiguousarray(Radius2),
        XiEtaZetaBasis=np.ascontiguousarray(XEZ), slenXiEtaZetaBasis=np.ascontiguousarray(slen),
        xGP=xGP, wGP=wGP, wBary=wBary, Elem_xGP=np.ascontiguousarray(Elem_xGP), sJ=np.ascontiguousarray(sJ),
        nBCs=len(bcs), bc_kind=bc_kind, bc_alpha=bc_alpha, PeriodicVectors=PV,
        Periodic_nNodes=Periodic_nNodes, Periodic_offsetNode=Periodic_offset, Periodic_Nodes=Periodic_Nodes,
        NodeVolume=NodeVolume, xyz_min=coords.min(axis=0), xyz_max=coords.max(axis=0),
        unique_coords=coords, elem_nodes=elem_nodes.astype(np.int32))


def from_hopr_arrays(ElemInfo, SideInfo, NodeCoords, GlobalNodeIDs, BCType, BCNames, N, part_bc=None,
                     tracking=TRIATRACKING) -> ParticleMesh:
    """Particle-mesh tables from the datasets of a HOPR mesh file (mesh/mesh_readin.f90, particle_mesh_readin.f90): elements,
    sides, side ids / flips and unique node ids are used as stored, so that element order, triangle diagonals and the owner
    of planar faces are those of the reference run on the same file.

    Hexahedra with 8 nodes (NGeo = 1, element types 108 / 118) without mortars.  BCType rows (type, curve, state, alpha):
    type 1 is periodic with alpha = +-vector id; every other boundary takes its particle condition from
    part_bc[name] (BC_OPEN / BC_REFLECTIVE, the Part-Boundary<n>-Condition of parameter.ini; default BC_OPEN).
    The periodic vectors are the offsets between paired periodic sides (GetPeriodicVectors)."""
    EI = np.asarray(ElemInfo, dtype=np.int64)
    SI = np.asarray(SideInfo, dtype=np.int64)
    NC = np.asarray(NodeCoords, dtype=np.float64)
    G = np.asarray(GlobalNodeIDs, dtype=np.int64)
    nE = EI.shape[0]
    if np.any((EI[:, 0] != 108) & (EI[:, 0] != 118)):
        raise ValueError("only 8-node hexahedra (element types 108, 118) are supported")
    if np.any(EI[:, 3] - EI[:, 2] != 6) or np.any(EI[:, 5] - EI[:, 4] != 8):
        raise ValueError("elements must have 6 sides and 8 nodes (NGeo = 1, no mortars)")
    if EI[0, 2] != 0 or np.any(EI[1:, 2] != EI[:-1, 3]) or EI[0, 4] != 0 or np.any(EI[1:, 4] != EI[:-1, 5]):
        raise ValueError("ElemInfo offsets are not contiguous")
    if SI.shape != (6 * nE, 5) or NC.shape != (8 * nE, 3) or G.shape != (8 * nE,):
        raise ValueError("SideInfo / NodeCoords / GlobalNodeIDs do not match ElemInfo")
    uniq, first, inv = np.unique(G, return_index=True, return_inverse=True)
    coords = NC[first]
    if np.abs(coords[inv] - NC).max() > 1e-12 * max(1.0, np.abs(NC).max()):
        raise ValueError("nodes with the same GlobalNodeID have different coordinates")
    tens = inv.reshape(nE, 8)                                  # unique id per element node, tensor order
    elem_nodes = tens[:, _CNS0]                                # CGNS corner order
    names = [(b.decode() if isinstance(b, bytes) else str(b)).strip() for b in BCNames]
    BT = np.asarray(BCType, dtype=np.int64).reshape(-1, 4)
    part_bc = part_bc or {}
    bcs = []
    for nme, row in zip(names, BT):
        bcs.append((BC_PERIODIC, int(row[3])) if row[0] == 1 else (part_bc.get(nme, BC_OPEN), 0))
    # periodic vectors from the paired sides
    nPV = int(np.abs(BT[BT[:, 0] == 1, 3]).max()) if np.any(BT[:, 0] == 1) else 0
    PV = np.zeros((nPV, 3))
    if nPV:
        fn = elem_nodes[:, _NODEMAP_CGNS0].reshape(6 * nE, 4)
        cen = coords[fn].mean(axis=1)
        alpha = np.where(SI[:, 4] > 0, BT[np.maximum(SI[:, 4], 1) - 1, 3] * (BT[np.maximum(SI[:, 4], 1) - 1, 0] == 1), 0)
        for k in range(1, nPV + 1):
            sd = np.nonzero(alpha == k)[0]
            if sd.size == 0:
                raise ValueError("periodic vector %d has no side with BC_ALPHA=+%d" % (k, k))
            nb = 6 * (SI[sd, 2] - 1) + (SI[sd, 3] // 10 - 1)
            v = cen[nb] - cen[sd]
            if np.abs(v - v[0]).max() > 1e-10 * max(1.0, np.abs(v).max()):
                raise ValueError("periodic sides of vector %d are not congruent" % k)
            # exact offset between stored node coordinates (the centroid means carry rounding fuzz): the partner-side node
            # that matches the first node of the side
            cand = coords[fn[nb[0]]] - coords[fn[sd[0], 0]]
            PV[k - 1] = cand[np.argmin(np.linalg.norm(cand - v[0], axis=1))]
    return build_mesh(coords, elem_nodes, N, bcs, None, periodic_vectors=PV, tracking=tracking, hopr_sides=SI)


def from_hopr_file(path, N, part_bc=None, tracking=TRIATRACKING) -> ParticleMesh:
    """ParticleMesh from a HOPR *_mesh.h5 (read with the built-in minimal HDF5 reader, h5mini.py)."""
    from .h5mini import H5File
    f = H5File(path)
    return from_hopr_arrays(f.read("ElemInfo"), f.read("SideInfo"), f.read("NodeCoords"), f.read("GlobalNodeIDs"),
                            f.read("BCType"), f.read("BCNames"), N, part_bc=part_bc, tracking=tracking)


def structured_connectivity(nx, ny, nz):
    """Unique-node ids (0-based, x fastest) of every element of an nx*ny*nz block, CGNS corner order."""
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    i = i.transpose(2, 1, 0).reshape(-1)   # element order: i fastest, then j, then k
    j = j.transpose(2, 1, 0).reshape(-1)
    k = k.transpose(2, 1, 0).reshape(-1)
    en = np.empty((nx * ny * nz, 8), dtype=np.int64)
    for c in range(8):
        a, b, cc = _CGNS_IJK[c]
        en[:, c] = (i + a) + (nx + 1) * ((j + b) + (ny + 1) * (k + cc))
    return en


def box_mesh(xmin, xmax, nelems, N, periodic=(True, True, True), wall_kind=BC_OPEN,
             tracking=TRIATRACKING, deform=None) -> ParticleMesh:
    """Cartesian hopr-style box (`Corner` = box, `nElems` = nelems), optionally deformed.

    periodic[d]   periodic in direction d (BC pair with BC_ALPHA = +-(index of its periodic vector))
    deform        optional callable(coords (n,3)) -> coords, applied to the unique grid nodes
                  (must keep opposite periodic faces congruent)
    BCIDs: 1..6 = x-, x+, y-, y+, z-, z+.
    """
    xmin = np.asarray(xmin, dtype=np.float64)
    xmax = np.asarray(xmax, dtype=np.float64)
    nx, ny, nz = (int(v) for v in nelems)
    gx = xmin[0] + (xmax[0] - xmin[0]) * (np.arange(nx + 1) / nx)
    gy = xmin[1] + (xmax[1] - xmin[1]) * (np.arange(ny + 1) / ny)
    gz = xmin[2] + (xmax[2] - xmin[2]) * (np.arange(nz + 1) / nz)
    gx[-1], gy[-1], gz[-1] = xmax
    Z, Y, X = np.meshgrid(gz, gy, gx, indexing="ij")
    coords = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
    ideal = coords.copy()
    if deform is not None:
        coords = np.asarray(deform(coords), dtype=np.float64)
    en = structured_connectivity(nx, ny, nz)

    pvs, bcs = [], []
    L = xmax - xmin
    for d in range(3):
        if periodic[d]:
            v = np.zeros(3)
            v[d] = L[d]
            pvs.append(v)
            pid = len(pvs)
            bcs += [(BC_PERIODIC, pid), (BC_PERIODIC, -pid)]
        else:
            bcs += [(wall_kind, 0), (wall_kind, 0)]

    # boundary faces are identified on the undeformed grid
    tol = 1e-9 * np.linalg.norm(L)

    def bc_of_face(cent, face_ids):
        ic = ideal[face_ids].mean(axis=1)
        out = np.zeros(cent.shape[0], dtype=np.int32)
        for d in range(3):
            out[np.abs(ic[:, d] - xmin[d]) <= tol] = 2 * d + 1
            out[np.abs(ic[:, d] - xmax[d]) <= tol] = 2 * d + 2
        return out

    m = build_mesh(coords, en, N, bcs, bc_of_face, periodic_vectors=pvs, tracking=tracking)
    m.extra.update(dict(nelems=(nx, ny, nz), box=(xmin.copy(), xmax.copy()), cartesian=deform is None))
    return m


def cartesian_locate(mesh: ParticleMesh, pos):
    """1-based global element id of points in an undeformed box_mesh (harness helper)."""
    nx, ny, nz = mesh.extra["nelems"]
    xmin, xmax = mesh.extra["box"]
    h = (xmax - xmin) / np.array([nx, ny, nz])
    ijk = np.floor((pos - xmin) / h).astype(np.int64)
    ijk = np.clip(ijk, 0, np.array([nx, ny, nz]) - 1)
    return (1 + ijk[:, 0] + nx * (ijk[:, 1] + ny * ijk[:, 2])).astype(np.int32)


def partition(mesh: ParticleMesh, nprocs: int):
    """Equal split of the element range (loadbalance/loaddistribution.f90:362-369); fills ELEM_RANK."""
    nG = mesh.nElems
    off = np.array([(nG // nprocs) * p + min(p, nG % nprocs) for p in range(nprocs + 1)], dtype=np.int64)
    off[nprocs] = nG
    rank = np.searchsorted(off, np.arange(nG), side="right") - 1
    mesh.ElemInfo[:, 6] = rank.astype(np.int32)
    return off


def add_fibgm(mesh: ParticleMesh, deltas=None, factor=(1.0, 1.0, 1.0)):
    """Fast-init background mesh (particle_bgm.f90:327-339, 441-450, 517-522): GEO%FIBGMdeltas = Part-FIBGMdeltas /
    Part-FactorFIBGM, cells 1..ceil(L/delta) per direction, every element registered in the cells its bounding box
    overlaps (ascending element id inside a cell).  deltas defaults to the mean element bounding-box size."""
    NC = mesh.NodeCoords.reshape(mesh.nElems, 8, 3)
    lo, hi = NC.min(axis=1), NC.max(axis=1)
    if deltas is None:
        deltas = (hi - lo).mean(axis=0)
    deltas = np.asarray(deltas, dtype=np.float64) / np.asarray(factor, dtype=np.float64)
    L = mesh.xyz_max - mesh.xyz_min
    deltas = np.minimum(deltas, L)
    nmax = np.floor(L / deltas).astype(np.int64) + 1
    nmax = np.where(np.mod(L, deltas) != 0, nmax, nmax - 1)                       # BGMimaxglob
    cmin = np.maximum(np.floor((lo - mesh.xyz_min) / deltas), 0).astype(np.int64) + 1
    cmax = np.minimum(np.floor((hi - mesh.xyz_min) / deltas).astype(np.int64) + 1, nmax[None, :])
    ElemToBGM = np.stack([cmin[:, 0], cmax[:, 0], cmin[:, 1], cmax[:, 1], cmin[:, 2], cmax[:, 2]], axis=1).astype(np.int32)
    ni, nj, nk = (int(v) for v in nmax)
    span = (cmax - cmin + 1)
    tot = span.prod(axis=1)
    e_rep = np.repeat(np.arange(mesh.nElems), tot)
    loc = np.arange(tot.sum()) - np.repeat(np.cumsum(tot) - tot, tot)
    sx, sy = span[e_rep, 0], span[e_rep, 1]
    ci = cmin[e_rep, 0] + loc % sx
    cj = cmin[e_rep, 1] + (loc // sx) % sy
    ck = cmin[e_rep, 2] + loc // (sx * sy)
    cell = (ci - 1) + ni * ((cj - 1) + nj * (ck - 1))
    order = np.lexsort((e_rep, cell))
    cnt = np.bincount(cell, minlength=ni * nj * nk)
    mesh.extra["FIBGM"] = dict(deltas=deltas, min=(1, 1, 1), max=(ni, nj, nk),
                               nElems=cnt.reshape(nk, nj, ni).astype(np.int32),
                               offsetElem=(np.cumsum(cnt) - cnt).reshape(nk, nj, ni).astype(np.int32),
                               Element=(e_rep[order] + 1).astype(np.int32))
    mesh.extra["ElemToBGM"] = np.ascontiguousarray(ElemToBGM)
    rad = np.zeros(mesh.nElems)
    for n in range(8):
        v = NC[:, n, :] - mesh.ElemBaryNGeo
        rad = np.maximum(rad, np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2]))
    mesh.extra["ElemRadiusNGeo"] = rad
    return mesh


def shape_function_setup(mesh: ParticleMesh, params, r_sf, alpha_sf, dim_sf=3, dim_sf_dir=1, sfDepo3D=True):
    """InitShapeFunctionDimensionalty (pic_depo_shapefunction_tools.f90:1131-1294): w_sf and dimFactorSF for the
    fixed-radius shape functions; fills the shape-function fields of `params`."""
    import math
    ext = mesh.xyz_max - mesh.xyz_min
    dimFactorSF = 1.0
    if dim_sf == 1:
        o = [d for d in range(3) if d != dim_sf_dir - 1]
        dimFactorSF = ext[o[0]] * ext[o[1]]
        w = math.gamma(float(alpha_sf) + 1.5) / (math.sqrt(math.pi) * r_sf * math.gamma(float(alpha_sf + 1)))
        w_sf = w / dimFactorSF if sfDepo3D else w
        if sfDepo3D:
            w_sf = math.gamma(float(alpha_sf) + 1.5) / (math.sqrt(math.pi) * r_sf * math.gamma(float(alpha_sf + 1)) * dimFactorSF)
    elif dim_sf == 2:
        dimFactorSF = ext[dim_sf_dir - 1]
        r2 = r_sf * r_sf
        w_sf = (float(alpha_sf) + 1.0) / (math.pi * r2 * dimFactorSF) if sfDepo3D else (float(alpha_sf) + 1.0) / (math.pi * r2)
    else:
        beta = math.gamma(1.5) * math.gamma(float(alpha_sf) + 1.0) / math.gamma(1.5 + float(alpha_sf) + 1.0)
        w_sf = 1.0 / (2.0 * beta * float(alpha_sf) + 2 * beta) * (float(alpha_sf) + 1.0) / (math.pi * (r_sf ** 3))
    params.r_sf, params.alpha_sf, params.dim_sf, params.dim_sf_dir = float(r_sf), int(alpha_sf), int(dim_sf), int(dim_sf_dir)
    params.sfDepo3D, params.w_sf, params.dimFactorSF = int(bool(sfDepo3D)), float(w_sf), float(dimFactorSF)
    return params


def add_refmapping_tables(mesh: ParticleMesh, bc_halo_eps=None, RefMappingEps=1e-4):
    """Side and BC-side tables of the RefMapping tracking (TrackingMethod = refmapping), straight-sided NGeo = 1 meshes.

    BaseVectors0/1/2            particle_mesh_build.f90:1833-1940 (Bezier control points of an NGeo = 1 side = its 4 corners)
    SideType/NormVec/Distance   particle_mesh_tools.f90:848-1056 (IdentifyElemAndSideType, linear-element branch)
    BCSideMetrics               particle_mesh_build.f90:1688-1830 (side origin = bilinear centre, radius = farthest corner)
    ElemToBCSides/SideBCMetrics particle_mesh_build.f90:645-1092  (BuildBCElemDistance; every SIDE_BCID>0 side is a BC side,
                                periodic ones included; an element lists its own BC sides and those within BC_halo_eps,
                                sorted by distance with the stable InsertionSort of utils.f90:52-101)
    ElemEpsOneCell              particle_mesh_build.f90:447-642   (1 + sqrt(3 scaleJ RefMappingEps))
    bc_halo_eps: halo_eps_velo*dt*SafetyFactor of the MPI build; None -> every element sees every BC side (the fullMesh
    branch, which is also what the single-rank build of the reference does).
    """
    if mesh.tracking != REFMAPPING:
        raise ValueError("build the mesh with tracking=REFMAPPING (ElemBaryNGeo = X(xi=0))")
    if "FIBGM" not in mesh.extra:
        add_fibgm(mesh)
    nE, nS = mesh.nElems, mesh.nSides
    P = mesh.NodeCoords[mesh.ElemSideNodeID.reshape(nS, 4)]             # p00, p10, p11, p01 = side nodes 1..4
    p00, p10, p11, p01 = P[:, 0], P[:, 1], P[:, 2], P[:, 3]
    BV0 = (+p00 + p10 + p01 + p11)
    BV1 = (-p00 + p10 - p01 + p11)
    BV2 = (-p00 - p10 + p01 + p11)
    BV3 = (+p00 - p10 - p01 + p11)
    cr = np.cross(BV1, BV2)
    nrm = cr / np.sqrt((cr * cr).sum(axis=1))[:, None]                   # CROSSNORM(v1,v2)
    centre = 0.25 * (p00 + p10 + p01 + p11)
    elem_of_side = np.arange(nS) // 6
    v2 = centre - mesh.ElemBaryNGeo[elem_of_side]
    S = mesh.SideInfo
    flip = np.where(S[:, 1] > 0, 0, S[:, 3] % 10)
    dot = (v2 * nrm).sum(axis=1)
    sign = np.where(flip == 0, np.where(dot < 0, -1.0, 1.0), np.where(dot > 0, -1.0, 1.0))
    SideNormVec = nrm * sign[:, None]
    SideDistance = (centre * SideNormVec).sum(axis=1)

    def unit(v):
        n = np.sqrt((v * v).sum(axis=1))
        return np.where(n[:, None] > 0, v / np.where(n > 0, n, 1.0)[:, None], 0.0)
    az = lambda a: np.abs(a) <= 2.22e-16                                  # ALMOSTZERO, piclas.h:83
    e1, e2, e3, e4 = unit(p01 - p00), unit(p10 - p00), unit(p11 - p01), unit(p11 - p10)
    rect = az((e1 * e2).sum(1)) & az((e1 * e3).sum(1)) & az((e4 * e2).sum(1)) & az((e4 * e3).sum(1))
    # planarity: the 4th corner lies in the plane of the other three
    planar = np.abs(((p11 - p00) * np.cross(p10 - p00, p01 - p00)).sum(1)) <= 1e-12 * np.abs(cr).max()
    SideType = np.where(planar & rect, 0, np.where(planar, 1, 2)).astype(np.int32)   # PLANAR_RECT / PLANAR_NONRECT / BILINEAR

    # BC sides (periodic included)
    bc_sides = np.nonzero(S[:, 4] > 0)[0]                                 # 0-based side index, ascending == BCSide order
    origin = centre[bc_sides]
    radius = np.sqrt(np.max(((P[bc_sides] - origin[:, None, :]) ** 2).sum(axis=2), axis=1))
    ElemRadius = mesh.extra["ElemRadiusNGeo"] if "ElemRadiusNGeo" in mesh.extra else np.sqrt(mesh.ElemRadius2NGeo)
    bary = mesh.ElemBaryNGeo
    diag = np.linalg.norm(mesh.xyz_max - mesh.xyz_min)
    full = bc_halo_eps is None or bc_halo_eps >= diag
    tree = None if full else cKDTree(origin)
    maxr = radius.max() if radius.size else 0.0
    ElemToBCSides = np.full((nE, 2), -1, dtype=np.int32)
    rows = []
    off = 0
    bc_elem = bc_sides // 6
    # bc_sides ascends, so the BC sides of element e are the contiguous range own_lo[e]:own_hi[e] (no scan per element);
    # with a halo distance all neighbourhood queries go through the tree in one call and elements far from every BC side
    # (the bulk of a large mesh) are skipped without touching the side arrays
    own_lo = np.searchsorted(bc_elem, np.arange(nE), side="left")
    own_hi = np.searchsorted(bc_elem, np.arange(nE), side="right")
    all_idx = np.arange(bc_sides.size)
    balls = None if full else tree.query_ball_point(bary, bc_halo_eps + ElemRadius + maxr)
    for e in range(nE):
        own = bc_sides[own_lo[e]:own_hi[e]]
        if full:
            other_idx = np.concatenate([all_idx[:own_lo[e]], all_idx[own_hi[e]:]])
        else:
            if not balls[e]:
                continue                                                  # own sides would be in the ball: no BC side in reach
            cand = np.array(sorted(balls[e]), dtype=np.int64)
            if cand.size:
                cand = cand[bc_elem[cand] != e]
                be = bc_elem[cand]
                ok = np.linalg.norm(bary[e] - bary[be], axis=1) <= bc_halo_eps + ElemRadius[e] + ElemRadius[be]
                ok &= np.linalg.norm(bary[e] - origin[cand], axis=1) <= bc_halo_eps + ElemRadius[e] + radius[cand]
                other_idx = cand[ok]
            else:
                other_idx = cand
        sides = np.concatenate([own, bc_sides[other_idx]]) if own.size or other_idx.size else np.zeros(0, dtype=np.int64)
        if sides.size == 0:
            continue
        # map side -> BC index for origin/radius
        bidx = np.searchsorted(bc_sides, sides)
        vec = bary[e][None, :] - origin[bidx]
        dist = np.sqrt((vec * vec).sum(axis=1)) - ElemRadius[e] - radius[bidx]
        order = np.argsort(dist, kind="stable")                          # InsertionSort is stable
        m = np.zeros((sides.size, 7))
        m[:, 0] = sides[order] + 1
        m[:, 1] = e + 1
        m[:, 2] = dist[order]
        m[:, 3] = radius[bidx][order]
        m[:, 4:7] = origin[bidx][order]
        rows.append(m)
        ElemToBCSides[e, 0] = sides.size                                  # ELEM_NBR_BCSIDES
        ElemToBCSides[e, 1] = off                                         # ELEM_FIRST_BCSIDE
        off += sides.size
    SideBCMetrics = np.concatenate(rows) if rows else np.zeros((0, 7))
    n1 = mesh.N + 1
    sJ = mesh.sJ.reshape(nE, -1)
    scaleJ = sJ.max(axis=1) / sJ.min(axis=1)
    mesh.extra.update(dict(BaseVectors0=np.ascontiguousarray(BV0), BaseVectors1=np.ascontiguousarray(BV1),
                           BaseVectors2=np.ascontiguousarray(BV2), BaseVectors3=np.ascontiguousarray(BV3), SideNormVec=np.ascontiguousarray(SideNormVec),
                           SideDistance=np.ascontiguousarray(SideDistance), SideType=SideType,
                           ElemToBCSides=ElemToBCSides, SideBCMetrics=np.ascontiguousarray(SideBCMetrics),
                           ElemEpsOneCell=1.0 + np.sqrt(3.0 * scaleJ * RefMappingEps),
                           BaseVectorsScale=0.25 * np.sqrt((cr * cr).sum(axis=1))))
    return mesh


def shape_function_adaptive_setup(mesh: ParticleMesh, params, alpha_sf, dim_sf=3, dim_sf_dir=1, sfDepo3D=True,
                                  SFAdaptiveDOF=None, smoothing=False):
    """shape_function_adaptive: per-element radius SFElemr2 (InitShapeFunctionAdaptive, pic_depo.f90:632-823) and the
    dimensional factors (InitShapeFunctionDimensionalty with r_sf_loc = 1).  Element neighbours are the elements sharing a
    unique node (BuildNodeNeighbourhood, particle_mesh_build.f90:1095-1382).  Supported on the device for periodic meshes
    or with smoothing (the paths that use the element radius of the particle's element)."""
    import math
    shape_function_setup(mesh, params, 1.0, alpha_sf, dim_sf=dim_sf, dim_sf_dir=dim_sf_dir, sfDepo3D=sfDepo3D)
    dimFactorSF = params.dimFactorSF
    Nmax = mesh.N
    if dim_sf == 1:
        default, DOFMax = 2.0 * (1. + 1.), 2.0 * (Nmax + 1.)
    elif dim_sf == 2:
        default, DOFMax = math.pi * (1. + 1.) ** 2, math.pi * (Nmax + 1.) ** 2
    else:
        default, DOFMax = (4. / 3.) * math.pi * (1. + 1.) ** 3, (4. / 3.) * math.pi * (Nmax + 1.) ** 3
    dof = default if SFAdaptiveDOF is None else float(SFAdaptiveDOF)
    if dof > DOFMax:
        raise ValueError("PIC-shapefunction-adaptive-DOF too large")
    scaling = dof / 2.0 if dim_sf == 1 else (math.sqrt(dof / math.pi) if dim_sf == 2 else (3. * dof / (4. * math.pi)) ** (1. / 3.))
    d1 = 1 if dim_sf_dir == 2 else 2
    d2 = 1 if dim_sf_dir == 3 else 3

    def sfnorm(v):
        if dim_sf == 1:
            return np.abs(v[..., dim_sf_dir - 1])
        if dim_sf == 2:
            return np.sqrt(v[..., d1 - 1] ** 2 + v[..., d2 - 1] ** 2)
        return np.sqrt((v * v).sum(axis=-1))

    def aer(a, b, tol):   # ALMOSTEQUALRELATIVE, piclas.h:79
        return np.abs(a - b) <= np.maximum(np.abs(a), np.abs(b)) * tol

    def measure(v1, v2):
        if dim_sf == 1:
            return ~aer(v1[..., dim_sf_dir - 1], v2[..., dim_sf_dir - 1], 1e-6)
        if dim_sf == 2:
            return ~(aer(v1[..., d1 - 1], v2[..., d1 - 1], 1e-6) & aer(v1[..., d2 - 1], v2[..., d2 - 1], 1e-6))
        return np.ones(v1.shape[:-1], dtype=bool)

    nE = mesh.nElems
    en = mesh.elem_nodes.astype(np.int64)                        # unique ids, CGNS order
    node2elem = [[] for _ in range(mesh.nUniqueNodes)]
    for e in range(nE):
        for u in en[e]:
            node2elem[u].append(e)
    X = mesh.unique_coords
    NC = mesh.NodeCoords.reshape(nE, 8, 3)
    r = np.full(nE, np.finfo(np.float64).max)
    for e in range(nE):
        mine = set(en[e].tolist())
        neigh = sorted({n for u in en[e] for n in node2elem[u]} - {e})
        done = False
        for nb in neigh:
            for ju in en[nb]:
                if ju in mine:
                    continue
                done = True
                v2 = X[ju]
                v1 = X[en[e]]
                ok = measure(v1, np.broadcast_to(v2, v1.shape))
                if ok.any():
                    r[e] = min(r[e], sfnorm(v1[ok] - v2).min())
        if not done:
            mid = NC[e].sum(axis=0) / 8.0
            r[e] = min(r[e], sfnorm(mid - NC[e]).min())
        bb = NC[e].max(axis=0) - NC[e].min(axis=0)
        vol = bb[0] * bb[1] * bb[2]
        cl = vol / dimFactorSF if dim_sf == 1 else (math.sqrt(vol / dimFactorSF) if dim_sf == 2 else vol ** (1. / 3.))
        if cl < r[e] or smoothing:
            r[e] = (r[e] + cl) / 2.0
        r[e] = r[e] * scaling / (mesh.N + 1.)
    mesh.extra["SFElemr2"] = np.ascontiguousarray(np.stack([r, r * r], axis=1))
    params.r_sf = float(r.max())
    return params
