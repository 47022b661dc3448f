"""CPU tests: the oracle against the reference's known-answer values and analytic self-checks,
host mesh tables, and the C ABI surface.  No GPU needed."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import cases
from oracle_lib import Oracle
from piclas_b200 import basis, hostmesh as hm
from piclas_b200.abi import Params, TIMEDISC_LEAPFROG

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_known_answers.json")))


# ---- basis ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 7])
def test_gauss_nodes_and_lagrange(N):
    x, w = basis.legendre_gauss_nodes_weights(N)
    xr, wr = np.polynomial.legendre.leggauss(N + 1)
    assert np.abs(x - xr).max() < 2e-15 and np.abs(w - wr).max() < 4e-15
    wb = basis.barycentric_weights(x)
    for xi in (-1.0, -0.3, 0.0, 0.77, 1.0):
        L = basis.lagrange_polys(xi, x, wb)
        assert abs(L.sum() - 1.0) < 1e-14
        assert abs((L * x ** N).sum() - xi ** N) < 1e-13         # exact for degree <= N
    L = basis.lagrange_polys(x[1], x, wb)                         # node hit -> exact unit vector (basis.f90:1254)
    assert L[1] == 1.0 and L.sum() == 1.0


def test_oracle_lagrange_matches_host_and_hits_nodes():
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (2, 2, 2), 3)
    o = Oracle(mesh, Params())
    for xi in np.linspace(-1, 1, 41):
        assert np.array_equal(o.lagrange(xi, mesh.xGP, mesh.wBary), basis.lagrange_polys(xi, mesh.xGP, mesh.wBary))
    # ALMOSTEQUAL_UNITY: within one epsilon relative of a node -> unit vector; absolute 2*eps test at zero
    eps = np.finfo(float).eps
    L = o.lagrange(mesh.xGP[2] * (1 + 0.5 * eps), mesh.xGP, mesh.wBary)
    assert L[2] == 1.0 and L.sum() == 1.0
    x5, _ = basis.legendre_gauss_nodes_weights(4)
    L = o.lagrange(1.5 * eps, x5, basis.barycentric_weights(x5))
    assert L[2] == 1.0


# ---- host mesh tables ---------------------------------------------------------------------------------------------
def test_mesh_tables_are_consistent():
    lo, hi = [-1, -1, -1], [1, 1, 1]
    for deform in (None, cases.wavy(0.05, lo, hi)):
        m = hm.box_mesh(lo, hi, (4, 3, 5), 2, deform=deform)
        S = m.SideInfo
        nb_elem, nb_side = S[:, 2], S[:, 7]
        assert (nb_elem > 0).all()                                      # fully periodic: every side has a neighbour
        # neighbour relation is symmetric and master/slave alternate
        back = S[nb_side - 1, 7]
        assert np.array_equal(S[nb_side - 1, 5], nb_elem)
        assert np.array_equal(np.abs(S[nb_side - 1, 1]), np.abs(S[:, 1]))
        inner = S[:, 4] == 0
        assert np.array_equal(back[inner], np.arange(1, m.nSides + 1)[inner])
        # both elements see the same first node and the same diagonal (node 1 -> node 3) on a shared inner face
        P = m.NodeCoords[m.ElemSideNodeID.reshape(-1, 4)]
        Pn = P[nb_side - 1]
        assert np.abs(P[inner, 0] - Pn[inner, 0]).max() < 1e-14
        assert np.abs(P[inner, 2] - Pn[inner, 2]).max() < 1e-14
        if deform is None:
            # planar faces: exactly one of the two elements owns the face (ConcaveElemSide tie break,
            # particle_mesh_tools.f90:1926-1928) unless an element is its own neighbour
            c = m.ConcaveElemSide.reshape(-1)
            assert ((c + c[nb_side - 1])[inner] == 1).all()
        # node volumes: every member of a periodic class carries the class total, classes sum to the box volume
        canon = np.arange(m.nUniqueNodes)
        for n in range(m.nUniqueNodes):
            k, o = m.Periodic_nNodes[n], m.Periodic_offsetNode[n]
            if k:
                canon[n] = min(n, (m.Periodic_Nodes[o:o + k] - 1).min())
        vol = sum(m.NodeVolume[n] for n in np.unique(canon))
        assert abs(vol - 8.0) < 1e-12


def test_partition_is_the_reference_split():
    m = hm.box_mesh([0, 0, 0], [1, 1, 1], (5, 3, 2), 1)
    off = hm.partition(m, 4)                                            # loaddistribution.f90:362-369
    assert list(off) == [0, 8, 16, 23, 30]
    assert np.array_equal(np.bincount(m.ElemInfo[:, 6]), [8, 8, 7, 7])


# ---- known answers of the reference's regression checks -------------------------------------------------------------
def test_plasma_ball_cvwm_deposited_charge():
    k = GOLD["NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean"]
    mesh, prm, PS, spec, elem = cases.plasma_ball_cvwm()
    o = Oracle(mesh, prm)
    PSrc, NS = o.deposit(PS, spec, elem, np.ones(len(spec), dtype=np.int32))
    assert abs(o.deposited_charge(PSrc) - k["charge"]) <= k["abs_tol"]
    # also after a few pushes through the periodic boundaries (charge is conserved by CVWM on any position set)
    inside = np.ones(len(spec), dtype=np.int32)
    isnew = np.zeros(len(spec), dtype=np.int32)
    PS[:, 3:] = np.random.default_rng(0).normal(0, 0.3, (len(spec), 3))
    E = np.zeros((mesh.nElems, 2, 2, 2, 3))
    for _ in range(6):
        assert o.push_track(1.0, PS, spec, elem, inside, isnew, E)[0] == 0
    assert np.abs(PS[:, :3]).max() <= 1.0
    PSrc, NS = o.deposit(PS, spec, elem, inside)
    assert abs(o.deposited_charge(PSrc) - k["charge"]) <= k["abs_tol"]


def test_charge_conserved_on_deformed_mesh():
    lo, hi = [-1, -1, -1], [1, 1, 1]
    mesh = hm.box_mesh(lo, hi, (3, 3, 3), 2, deform=cases.wavy(0.08, lo, hi))
    prm = cases.electron_params(MacroParticleFactor=(1e12,))
    PS, spec = cases.uniform_plasma(mesh, 4000, seed=4)
    o = Oracle(mesh, prm)
    elem = o.locate(PS[:, :3])
    assert (elem > 0).all()
    PSrc, NS = o.deposit(PS, spec, elem, np.ones(4000, dtype=np.int32))
    q = 4000 * 1e12 * (-cases.QE)
    # integrating the deposited density with the nodal quadrature reproduces the charge up to the quadrature error of
    # the non-constant Jacobian (the reference tolerates 1e-3 on its deformed case, analyze.ini of ..._save_CVWM)
    assert abs(o.deposited_charge(PSrc) - q) <= 1e-3 * abs(q)


# ---- analytic self-checks ---------------------------------------------------------------------------------------------
def test_newton_on_cartesian_and_deformed_elements():
    mesh = hm.box_mesh([0, 0, 0], [2, 1, 1], (4, 2, 2), 2)
    o = Oracle(mesh, Params())
    rng = np.random.default_rng(1)
    x = rng.random((2000, 3)) * [2, 1, 1]
    el = hm.cartesian_locate(mesh, x)
    xi, suc, bad = o.position_in_ref_elem(x, el)
    h = np.array([0.5, 0.5, 0.5])
    exact = 2 * (x / h - np.floor(x / h)) - 1
    assert bad == 0 and suc.all() and np.abs(xi - exact).max() < 1e-14
    lo, hi = [-1, -1, -1], [1, 1, 1]
    md = hm.box_mesh(lo, hi, (3, 3, 3), 2, deform=cases.wavy(0.08, lo, hi))
    od = Oracle(md, Params())
    x = rng.uniform(-1, 1, (2000, 3))
    el = od.locate(x)
    xi, suc, bad = od.position_in_ref_elem(x, el)
    # TriaTracking's element (planar triangles) and the trilinear map differ near non-planar faces: |xi| may exceed 1 slightly
    assert suc.all() and np.abs(xi).max() <= 1.2
    # X(xi) reproduces x to the Newton tolerance RefMappingEps = 1e-4 on |delta xi|^2 (eval_xyz.f90:359)
    w = 0.5 * np.stack([1 - xi, 1 + xi], axis=-1)                       # (n,3,2)
    X = np.zeros_like(x)
    for k in range(2):
        for j in range(2):
            for i in range(2):
                X += md.XCL_NGeo[el - 1, k, j, i] * (w[:, 0, i] * w[:, 1, j] * w[:, 2, k])[:, None]
    assert np.abs(X - x).max() < 2e-3


def test_interpolation_is_exact_for_polynomials():
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (3, 3, 3), 3)
    o = Oracle(mesh, cases.electron_params())
    PS, spec = cases.uniform_plasma(mesh, 3000, seed=9)
    el = hm.cartesian_locate(mesh, PS[:, :3])
    X = mesh.Elem_xGP
    E = np.zeros(X.shape)
    E[..., 0] = X[..., 0] ** 3 - X[..., 1] * X[..., 2]
    E[..., 1] = X[..., 1] ** 2 * X[..., 0]
    E[..., 2] = 1 + X[..., 2] ** 3
    F = o.interpolate(PS, el, E)
    x = PS[:, :3]
    assert np.abs(F[:, 0] - (x[:, 0] ** 3 - x[:, 1] * x[:, 2])).max() < 1e-13
    assert np.abs(F[:, 1] - x[:, 1] ** 2 * x[:, 0]).max() < 1e-13
    assert np.abs(F[:, 2] - (1 + x[:, 2] ** 3)).max() < 1e-13
    assert not F[:, 3:].any()


def test_tracking_agrees_with_cartesian_floor_and_wraps_periodically():
    mesh = hm.box_mesh([-1, -1, -1], [1, 1, 1], (6, 5, 4), 1)
    o = Oracle(mesh, cases.electron_params())
    n = 5000
    PS, spec = cases.uniform_plasma(mesh, n, seed=3, vth_cells=1.7, dt=1.0)   # crosses several elements per step
    el = hm.cartesian_locate(mesh, PS[:, :3])
    inside = np.ones(n, dtype=np.int32)
    isnew = np.zeros(n, dtype=np.int32)
    E = np.zeros((mesh.nElems, 2, 2, 2, 3))
    for _ in range(5):
        assert o.push_track(1.0, PS, spec, el, inside, isnew, E)[0] == 0
        assert np.abs(PS[:, :3]).max() <= 1.0
        assert np.array_equal(el, hm.cartesian_locate(mesh, PS[:, :3]))


def test_boris_with_zero_B_equals_relativistic_leapfrog_and_leapfrog_limit():
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (2, 2, 2), 2)
    PS, spec = cases.uniform_plasma(mesh, 500, seed=6, vth_cells=0.2, dt=1e-6)
    el = hm.cartesian_locate(mesh, PS[:, :3])
    E = cases.smooth_field(mesh, 1e-3)
    inside = np.ones(500, dtype=np.int32)
    out = {}
    for name, td in (("boris", 508), ("leap", TIMEDISC_LEAPFROG)):
        o = Oracle(mesh, cases.electron_params(TimeDiscMethod=td))
        P, e = PS.copy(), el.copy()
        o.push_track(1e-6, P, spec, e, inside.copy(), np.zeros(500, dtype=np.int32), E)
        out[name] = P
    # v << c: the relativistic Boris step with B = 0 reduces to the leapfrog kick up to O(v^2/c^2)
    assert np.abs(out["boris"][:, 3:] - out["leap"][:, 3:]).max() / np.abs(out["leap"][:, 3:]).max() < 1e-5


def test_boris_rotation_in_a_uniform_magnetic_field():
    """Boris-Leapfrog with E = 0, B = B e_z (external field): the speed is conserved to rounding and the velocity turns by
    exactly q B dt / (gamma m) per step about B — the reference takes t = TAN(c_1 B / gamma), i.e. the exact half angle
    (timedisc_TimeStepPoissonByBorisLeapfrog.f90:153-185)."""
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (2, 2, 2), 1)
    B = 2e-4
    prm = cases.electron_params(externalField=(0.0, 0.0, 0.0, 0.0, 0.0, B), DoDeposition=0)
    o = Oracle(mesh, prm)
    n, dt = 200, 1e-9
    rng = np.random.default_rng(4)
    v0 = rng.normal(0, 3e6, (n, 3))
    PS = np.ascontiguousarray(np.concatenate([rng.uniform(0.3, 0.7, (n, 3)), v0], axis=1))
    spec = np.ones(n, dtype=np.int32)
    el = hm.cartesian_locate(mesh, PS[:, :3])
    inside, isnew = np.ones(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    E = np.zeros(mesh.Elem_xGP.shape)
    o.push_track(dt, PS, spec, el, inside, isnew, E)
    v1 = PS[:, 3:]
    assert np.abs(np.linalg.norm(v1, axis=1) / np.linalg.norm(v0, axis=1) - 1.0).max() < 1e-14
    assert np.abs(v1[:, 2] - v0[:, 2]).max() <= 1e-15 * np.abs(v0).max() * 4             # parallel component untouched
    gamma = 1.0 / np.sqrt(1.0 - (v0 * v0).sum(1) * prm.c2_inv)
    ang = np.arctan2(v1[:, 1], v1[:, 0]) - np.arctan2(v0[:, 1], v0[:, 0])
    ang = (ang + np.pi) % (2 * np.pi) - np.pi
    expect = cases.QE * B * dt / (gamma * cases.ME)                                    # electrons turn counter-clockwise about +z
    assert np.abs(ang - expect).max() < 1e-13


def test_open_boundary_removes_and_counts_nothing_lost():
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (3, 3, 3), 1, periodic=(False, False, False))
    o = Oracle(mesh, cases.electron_params())
    n = 2000
    PS, spec = cases.uniform_plasma(mesh, n, seed=12, vth_cells=1.0, dt=1.0)
    el = hm.cartesian_locate(mesh, PS[:, :3])
    inside = np.ones(n, dtype=np.int32)
    lost, _, _ = o.push_track(1.0, PS, spec, el, inside, np.zeros(n, dtype=np.int32), np.zeros((27, 2, 2, 2, 3)))
    out = (PS[:, :3] < 0).any(axis=1) | (PS[:, :3] > 1).any(axis=1)
    assert lost == 0 and np.array_equal(inside == 0, out)


@pytest.mark.parametrize("tracking", ["tria", "ref"])
def test_reflective_walls_fold_the_free_flight(tracking):
    """Specular walls at rest (PerfectReflection): without a field the flight in a box is the folded straight line, speeds are
    conserved and nobody leaves; several reflections per step and corner hits included.  Both tracking methods."""
    ref = tracking == "ref"
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (3, 2, 4), 1, periodic=(False, False, False), wall_kind=hm.BC_REFLECTIVE,
                       tracking=hm.REFMAPPING if ref else hm.TRIATRACKING)
    if ref:
        hm.add_fibgm(mesh)
        hm.add_refmapping_tables(mesh)
    o = Oracle(mesh, cases.electron_params(DoInterpolation=0, DoDeposition=0, TrackingMethod=mesh.tracking))
    n = 3000
    rng = np.random.default_rng(21)
    x0 = rng.uniform(0.02, 0.98, (n, 3))
    v0 = rng.normal(0.0, 0.9, (n, 3))          # up to a few box lengths per step
    PS = np.ascontiguousarray(np.concatenate([x0, v0], axis=1))
    spec = np.ones(n, dtype=np.int32)
    el = hm.cartesian_locate(mesh, x0)
    inside = np.ones(n, dtype=np.int32)
    E = np.zeros((mesh.nElems, 2, 2, 2, 3))
    xi = o.position_in_ref_elem(x0, el, force=False)[0] if ref else None
    t = 0.0
    for _ in range(3):
        lost, _, _ = o.push_track(1.0, PS, spec, el, inside, np.zeros(n, dtype=np.int32), E, PartPosRef=xi)
        t += 1.0
        assert lost == 0 and inside.all()
        y = np.mod(x0 + v0 * t, 2.0)
        fold = np.where(y > 1.0, 2.0 - y, y)
        assert np.abs(PS[:, :3] - fold).max() < 1e-12
        assert np.abs(np.abs(PS[:, 3:]) - np.abs(v0)).max() < 1e-14
        assert np.array_equal(el, hm.cartesian_locate(mesh, PS[:, :3]))


def test_oracle_reproduces_the_references_deposited_source_per_dof():
    """The one per-DOF vector the reference's regression checks hold for this path: its restart state of
    NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean (h5diff reference for DG_Source) with the 3333 particles and the charge
    density the reference deposited from them.  The oracle must reproduce all 8000 values."""
    mesh, prm, PS, spec, el, rho_ref, _ = cases.reference_plasma_ball()
    o = Oracle(mesh, prm)
    assert np.array_equal(el, o.locate(PS[:, :3]))
    src, _ = o.deposit(PS, spec, el, np.ones(len(spec), dtype=np.int32))
    assert not src[..., :3].any()
    assert np.abs(src[..., 3] - rho_ref).max() <= 1e-13 * np.abs(rho_ref).max()       # measured 3e-15
    assert abs(o.deposited_charge(src) - GOLD["NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean"]["charge"]) <= 5e-13


@pytest.mark.skipif(not os.path.isdir("/root/reference/regressioncheck"), reason="reference tree not mounted")
def test_committed_fixtures_match_the_reference_files():
    """Where the reference tree is mounted (the build container): the committed golden fixtures are exactly what the built-in HDF5
    reader extracts from the reference's files today."""
    from piclas_b200.h5mini import H5File
    d = "/root/reference/regressioncheck/NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean/"
    st = H5File(d + "plasma_wave_State_000.00000000000000000_restart.h5")
    g = np.load(os.path.join(ROOT, "tests", "golden", "plasma_ball_cvwm_reference.npz"))
    assert np.array_equal(st.read("PartData"), g["PartData"])
    assert np.array_equal(st.read("DG_Source")[..., 3], g["DG_Source_charge"])
    assert np.array_equal(st.read("PartInt"), g["PartInt"])
    gm = np.load(os.path.join(ROOT, "tests", "golden", "hopr_meshes.npz"))
    me = H5File(d + "Box_mesh.h5")
    for ds in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType"):
        assert np.array_equal(me.read(ds), gm["box_" + ds])
    tw = H5File("/root/reference/tutorials/pic-poisson-plasma-wave/plasma_wave_mesh.h5")
    assert np.array_equal(tw.read("NodeCoords"), gm["plasma_wave_NodeCoords"])


def _hopr(tag, N, **kw):
    g = np.load(os.path.join(ROOT, "tests", "golden", "hopr_meshes.npz"))
    return hm.from_hopr_arrays(*[g[tag + "_" + d] for d in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType", "BCNames")],
                               N, **kw)


def test_hopr_mesh_file_in_its_own_element_order_reproduces_dg_source():
    """hostmesh.from_hopr_arrays on the datasets of the regression check's Box_mesh.h5 (element order along HOPR's space-filling
    curve, its side ids and flips): the reference's DG_Source is reproduced in file order, and the tables agree with the
    generated box mesh element by element."""
    mesh = _hopr("box", 1)
    assert mesh.nElems == 1000 and mesh.nUniqueNodes == 1331
    assert np.abs(mesh.PeriodicVectors - 2.0 * np.eye(3)).max() < 1e-14      # offsets of the stored nodes (HOPR's -1 + 10 * 0.2)
    g = np.load(os.path.join(ROOT, "tests", "golden", "plasma_ball_cvwm_reference.npz"))
    PS = np.ascontiguousarray(g["PartData"][:, :6])
    spec = g["PartData"][:, 6].astype(np.int32)
    prm = Params(ChargeIC=(1.60217653e-5, -cases.QE), MassIC=(1.0, cases.ME), MacroParticleFactor=(200.0, 200.0))
    o = Oracle(mesh, prm)
    el = o.locate(PS[:, :3])
    # PartInt of the state file: the particles are stored element by element in the file's element order
    pi = g["PartInt"].T
    assert np.array_equal(np.repeat(np.arange(1, 1001), pi[:, 1] - pi[:, 0]), el)
    src, _ = o.deposit(PS, spec, el, np.ones(len(spec), dtype=np.int32))
    ref = g["DG_Source_charge"]
    assert np.abs(src[..., 3] - ref).max() <= 1e-13 * np.abs(ref).max()
    # same geometry as the generated mesh, element by element
    box = hm.box_mesh([-1, -1, -1], [1, 1, 1], (10, 10, 10), 1)
    ijk = np.floor((mesh.ElemBaryNGeo + 1.0) / 0.2).astype(int)
    ours = ijk[:, 0] + 10 * (ijk[:, 1] + 10 * ijk[:, 2])
    assert np.abs(mesh.XCL_NGeo - box.XCL_NGeo[ours]).max() < 1e-14
    nb_file = mesh.SideInfo[:, 2].reshape(1000, 6)
    nb_box = box.SideInfo[:, 2].reshape(1000, 6)[ours]
    assert np.array_equal(ours[nb_file - 1] + 1, nb_box)          # same neighbour through every local side
    nv = np.zeros(box.nUniqueNodes)
    key = lambda c: np.round((c + 1.0) / 0.2).astype(int) @ np.array([1, 11, 121])
    nv[key(box.unique_coords)] = box.NodeVolume
    assert np.abs(mesh.NodeVolume - nv[key(mesh.unique_coords)]).max() < 1e-15


@pytest.mark.parametrize("tag", ["box", "deformed", "plasma_wave"])
def test_generated_meshes_use_hoprs_flip_convention(tag):
    """hostmesh.build_mesh derives master / slave flips geometrically ("the slave's node that coincides with the master's first
    node"); on the reference's mesh files that rule must reproduce the flip digit HOPR stored for every slave side."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "hopr_meshes.npz"))
    EI, SI, NC, G = [g[tag + "_" + d] for d in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs")]
    nE = len(EI)
    _, first, inv = np.unique(G, return_index=True, return_inverse=True)
    coords = NC[first]
    fn = inv.reshape(nE, 8)[:, hm._CNS0][:, hm._NODEMAP_CGNS0].reshape(6 * nE, 4)
    S = SI.astype(np.int64)
    slaves = np.nonzero((S[:, 1] < 0) & (S[:, 2] > 0))[0]
    nb = 6 * (S[slaves, 2] - 1) + (S[slaves, 3] // 10 - 1)
    cen = coords[fn].mean(axis=1)
    shift = cen[slaves] - cen[nb]                       # zero for inner sides, the periodic vector otherwise
    shift[np.abs(shift) < 1e-9] = 0.0
    dist = np.linalg.norm(coords[fn[slaves]] - (coords[fn[nb, 0]] + shift)[:, None, :], axis=2)
    assert len(slaves) > 0 and np.array_equal(np.argmin(dist, axis=1) + 1, S[slaves, 3] % 10)


def test_twisted_mesh_from_the_hopr_file_known_answer():
    """Box_deformed_mesh.h5 as written by HOPR (its master / slave choice fixes the diagonal of the twisted interface; BC 7 is the
    inner dielectric boundary of that case): deposited charge within the regression check's 1e-3."""
    mesh = _hopr("deformed", 1, part_bc={n: hm.BC_REFLECTIVE for n in
                                          ("BC_x+", "BC_x-", "BC_y+", "BC_y-", "BC_z+", "BC_z-", "BC_DIELECTRIC")})
    assert mesh.nElems == 2 and int(mesh.SideInfo[2, 4]) == 7 and int(mesh.SideInfo[2, 1]) == -3
    _, prm, PS, spec = cases.plasma_ball_two_elements(True)
    o = Oracle(mesh, prm)
    el = o.locate(PS[:, :3])
    assert (el > 0).all()
    src, _ = o.deposit(PS, spec, el, np.ones(len(spec), dtype=np.int32))
    k = GOLD["NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean_save_CVWM"]
    assert abs(o.deposited_charge(src) - k["charge_deformed"]) <= k["abs_tol_deformed"]


def test_two_element_mesh_matches_the_references_mesh_file():
    """Corner nodes of cases.plasma_ball_two_elements(True) against Box_deformed_mesh.h5 of the regression check."""
    mesh, _, _, _ = cases.plasma_ball_two_elements(True)
    g = np.load(os.path.join(ROOT, "tests", "golden", "plasma_ball_cvwm_reference.npz"))
    ref = g["deformed_mesh_NodeCoords"]                               # (2, 8, 3), tensor order
    ours = np.asarray(mesh.NodeCoords).reshape(2, 8, 3)
    assert np.abs(ours - ref).max() <= 1e-15          # same corners in the same tensor order (HOPR's 0.4 is off by one ulp)


@pytest.mark.parametrize("deformed", [False, True])
def test_plasma_ball_two_elements_known_answer(deformed):
    """NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean_save_CVWM (analyze.ini: Charge 10.68010874898 at 5e-13 on the Cartesian
    and 1e-3 on the deformed two-element mesh).  The twisted interface makes Newton leave [-1,1]^3 for particles that the
    triangle test assigns to the convex element: the SucRefPos=F inverse-distance branch (pic_depo_method.f90:512-538)."""
    mesh, prm, PS, spec = cases.plasma_ball_two_elements(deformed)
    o = Oracle(mesh, prm)
    el = o.locate(PS[:, :3])
    assert (el > 0).all()
    _, suc, _ = o.position_in_ref_elem(PS[:, :3], el, force=True)
    assert ((np.asarray(suc) == 0).sum() > 50) == deformed          # the fallback is exercised on the deformed mesh only
    src, _ = o.deposit(PS, spec, el, np.ones(len(spec), dtype=np.int32))
    k = GOLD["NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean_save_CVWM"]
    ref, tol = (k["charge_deformed"], k["abs_tol_deformed"]) if deformed else (k["charge_cartesian"], k["abs_tol_cartesian"])
    assert abs(o.deposited_charge(src) - ref) <= tol


# ---- C ABI surface ---------------------------------------------------------------------------------------------------------
def test_library_exports_every_symbol_of_the_header():
    from piclas_b200 import build, lib
    build.build_cuda()
    hdr = open(os.path.join(ROOT, "include", "piclas_gpu.h")).read()
    names = sorted(set(re.findall(r"\b(piclas_gpu_\w+)\s*\(", hdr)))
    assert len(names) >= 14
    so = C.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(so, n), n
    assert sorted(names) == sorted(lib.EXPORTS)


def test_product_path_has_no_cpu_fallback():
    # the package never imports or links the oracle; without a CUDA device the ABI reports an error instead of computing
    import piclas_b200.particle_step as ps
    src = open(ps.__file__).read() + open(os.path.join(ROOT, "piclas_b200", "multi.py")).read()
    assert "oracle" not in src.lower().replace("the cpu (gloo) tests, which drive it with a cpu engine", "")
    # no source file of the package or of the boundary loads, links or names the checker's library; the only mentions allowed are
    # the build helper that compiles it (building the checker is not using it) and one docstring
    exts = (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".f90", ".inc")
    for base in (os.path.join(ROOT, "piclas_b200"), os.path.join(ROOT, "include")):
        for d, _, files in os.walk(base):
            for f in files:
                if not f.endswith(exts):
                    continue
                text = open(os.path.join(d, f), errors="ignore").read().lower()
                assert "liboracle" not in text and "oracle_lib" not in text and "piclas_oracle" not in text, f
                if f not in ("build.py", "hostmesh.py"):
                    assert "oracle" not in text, f
    import torch
    if not torch.cuda.is_available():
        mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (2, 2, 2), 1)
        with pytest.raises(ps.PiclasGpuError):
            ps.ParticleStep(mesh, Params())


# ---- shape function known answers (NIG_PIC_Deposition/Plasma_Ball_Shape-function-{x,y,z}Dir) --------------------------------
_SF_CASES = GOLD["NIG_PIC_Deposition/Plasma_Ball_Shape-function"]["cases"]


@pytest.mark.parametrize("kind", ["sf", "cc", "adaptive"])
@pytest.mark.parametrize("case", _SF_CASES, ids=[c["case"].replace(" ", "-") for c in _SF_CASES])
def test_oracle_shape_function_known_answers(case, kind):
    """All 18 reference values of the three regression directories: 1-D and 2-D shape functions in every direction,
    shape_function (5 % relative), shape_function_cc and shape_function_adaptive with smoothing (1e-9 absolute)."""
    from piclas_b200.abi import DEPO_SF, DEPO_SF_CC, DEPO_SF_ADAPTIVE
    k = GOLD["NIG_PIC_Deposition/Plasma_Ball_Shape-function"]
    mesh = hm.box_mesh([-5, -5, -5], [5, 5, 5], tuple(case["nelems"]), 3)
    hm.add_fibgm(mesh)
    n = 20000                                     # the charge is linear in the particle number: scale the known answer
    x = cases.sphere_points(np.random.default_rng(1), n, 0.5)
    PS = np.zeros((n, 6))
    PS[:, :3] = x
    prm = Params(ChargeIC=(1.60217653e-5,), MassIC=(1.0,), MacroParticleFactor=(200.0,),
                 DepositionType={"sf": DEPO_SF, "cc": DEPO_SF_CC, "adaptive": DEPO_SF_ADAPTIVE}[kind])
    if kind == "adaptive":
        hm.shape_function_adaptive_setup(mesh, prm, 2, dim_sf=case["dim"], dim_sf_dir=case["dir"], sfDepo3D=True, smoothing=True)
    else:
        hm.shape_function_setup(mesh, prm, 2.0, 2, dim_sf=case["dim"], dim_sf_dir=case["dir"], sfDepo3D=True)
    o = Oracle(mesh, prm)
    PSrc, _ = o.deposit(PS, np.ones(n, dtype=np.int32), hm.cartesian_locate(mesh, x), np.ones(n, dtype=np.int32))
    q = o.deposited_charge(PSrc) * (100000 / n)
    tol = k["tolerances"][kind]
    assert abs(q - case[kind]) <= (tol["value"] * case[kind] if tol["type"] == "relative" else tol["value"])


def _ref_points(mesh, n, rng):
    """Random points by element and reference position (trilinear map of the corner nodes)."""
    el = rng.integers(1, mesh.nElems + 1, n).astype(np.int32)
    xi = rng.uniform(-0.999, 0.999, (n, 3))
    w = lambda t: np.stack([(1 - t) / 2, (1 + t) / 2], -1)
    W = np.einsum("nk,nj,ni->nkji", w(xi[:, 2]), w(xi[:, 1]), w(xi[:, 0]))
    return el, np.einsum("nkji,nkjix->nx", W, mesh.XCL_NGeo[el - 1])


def test_refmapping_bilinear_and_nonrect_periodic_sides():
    """ComputeBiLinearIntersection (BILINEAR and PLANAR_NONRECT BC sides), the bilinear side normal and the
    LocateParticleInElement fallback: on a periodic mesh whose boundary is deformed nobody is lost, the positions equal the free
    flight up to whole periodic vectors, and the stored reference positions are those of the final element."""
    lo, hi = [0, 0, 0], [1, 1, 1]
    mesh = hm.box_mesh(lo, hi, (4, 4, 3), 2, tracking=hm.REFMAPPING, deform=cases.wavy_periodic(0.04, lo, hi))
    hm.add_fibgm(mesh)
    hm.add_refmapping_tables(mesh)
    bc_types = np.bincount(mesh.extra["SideType"][mesh.SideInfo[:, 4] > 0], minlength=3)
    assert bc_types[1] > 0 and bc_types[2] > 0
    o = Oracle(mesh, cases.electron_params(TrackingMethod=hm.REFMAPPING, DoDeposition=0))
    n, dt = 6000, 1e-8
    rng = np.random.default_rng(2)
    el, x = _ref_points(mesh, n, rng)
    v = rng.normal(0, 0.25 / dt, (n, 3))
    PS = np.ascontiguousarray(np.concatenate([x, v], axis=1))
    spec = np.ones(n, dtype=np.int32)
    xi, _, _ = o.position_in_ref_elem(PS[:, :3], el, force=False)
    inside, isnew = np.ones(n, dtype=np.int32), np.ones(n, dtype=np.int32)
    E = np.zeros(mesh.Elem_xGP.shape)
    free = x.copy()
    relocated = 0
    for it in range(8):
        lost, _, _ = o.push_track(dt, PS, spec, el, inside, isnew, E, PartPosRef=xi)
        free += v * dt
        assert lost == 0 and inside.all()
        d = PS[:, :3] - free
        assert np.abs(d - np.round(d)).max() < 1e-13                      # shifted by whole periodic vectors only
        chk, _, _ = o.position_in_ref_elem(PS[:, :3], el, force=False)
        assert np.array_equal(chk, xi) and np.abs(xi).max() <= mesh.extra["ElemEpsOneCell"].max()
        relocated += int(isnew.sum())
        isnew[:] = 0
    assert relocated >= 1                                                 # the fallback is exercised by the fastest particles


def test_refmapping_on_the_tutorial_mesh_file():
    """tutorials/pic-poisson-plasma-wave/plasma_wave_mesh.h5: HOPR's rounding leaves two boundary sides 8e-15 away from
    rectangular, which the reference's IdentifyElemAndSideType classifies PLANAR_NONRECT (ALMOSTZERO test); tracking through the
    bilinear intersection there must agree with the exactly rectangular generated mesh."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "hopr_meshes.npz"))
    mf = hm.from_hopr_arrays(*[g["plasma_wave_" + d] for d in ("ElemInfo", "SideInfo", "NodeCoords", "GlobalNodeIDs", "BCType", "BCNames")],
                             5, tracking=hm.REFMAPPING)
    ms = hm.box_mesh([0, 0, 0], [6.2831, 0.2, 0.2], (60, 1, 1), 5, tracking=hm.REFMAPPING)
    res = []
    for mesh in (mf, ms):
        hm.add_fibgm(mesh, deltas=(6.2831, 0.2, 0.2), factor=(60, 1, 1))
        hm.add_refmapping_tables(mesh)
        o = Oracle(mesh, cases.electron_params(TrackingMethod=hm.REFMAPPING, DoDeposition=0))
        n = 400
        xs = (np.arange(n) + 0.5) * 6.2831 / n
        PS = np.zeros((n, 6))
        PS[:, 0] = xs
        PS[:, 1:3] = 0.1
        PS[:, 3], PS[:, 4], PS[:, 5] = 3e6 * np.sin(xs), 2e6 * np.cos(3 * xs), -1e6 * np.sin(2 * xs)
        spec = np.ones(n, dtype=np.int32)
        el = o.locate(PS[:, :3])
        xi, _, _ = o.position_in_ref_elem(PS[:, :3], el, force=False)
        inside, isnew = np.ones(n, dtype=np.int32), np.ones(n, dtype=np.int32)
        E = np.zeros(mesh.Elem_xGP.shape)
        for _ in range(60):
            lost, _, _ = o.push_track(5e-10, PS, spec, el, inside, isnew, E, PartPosRef=xi)
            assert lost == 0
        res.append((PS.copy(), mesh.ElemBaryNGeo[el - 1, 0].copy()))
    assert np.bincount(mf.extra["SideType"])[1] == 2 and np.bincount(ms.extra["SideType"])[0] == 360
    assert np.abs(res[0][1] - res[1][1]).max() < 1e-12                      # same elements
    assert np.abs(res[0][0][:, :3] - res[1][0][:, :3]).max() < 1e-12


# ---- RefMapping tracking ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(6, 5, 4), (4, 4, 1), (12, 1, 1)])
def test_refmapping_agrees_with_triatracking_and_floor(shape):
    """On Cartesian periodic boxes both tracking methods must put every particle into the same element, including the
    1-element-thick tutorial meshes where every element is a BC element (ParticleBCTracking path)."""
    mt = hm.box_mesh([0, 0, 0], [1, 1, 1], shape, 2)
    mr = hm.box_mesh([0, 0, 0], [1, 1, 1], shape, 2, tracking=hm.REFMAPPING)
    hm.add_fibgm(mr)
    hm.add_refmapping_tables(mr)
    ot, orr = Oracle(mt, cases.electron_params()), Oracle(mr, cases.electron_params(TrackingMethod=hm.REFMAPPING))
    n, dt = 5000, 1e-8
    PS, spec = cases.uniform_plasma(mt, n, seed=3, vth_cells=0.6, dt=dt)
    el = hm.cartesian_locate(mt, PS[:, :3])
    E = cases.smooth_field(mt, 1e-4)
    xi, _, _ = orr.position_in_ref_elem(PS[:, :3], el, force=False)
    A = [PS.copy(), el.copy(), np.ones(n, dtype=np.int32), np.ones(n, dtype=np.int32)]
    B = [PS.copy(), el.copy(), np.ones(n, dtype=np.int32), np.ones(n, dtype=np.int32)]
    for _ in range(5):
        ot.push_track(dt, A[0], spec, A[1], A[2], A[3], E)
        orr.push_track(dt, B[0], spec, B[1], B[2], B[3], E, PartPosRef=xi)
        assert np.array_equal(A[1], B[1])
        assert np.array_equal(B[1], hm.cartesian_locate(mr, B[0][:, :3]))
        assert np.abs(A[0] - B[0]).max() <= 1e-12 * np.abs(A[0]).max()
        assert np.abs(xi).max() < 1.0
