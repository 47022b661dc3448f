"""The device math headers of the CUDA path (piclas_b200/csrc/math.cuh, fastmath.cuh) compiled for the host
(tests/device_math_host.cpp), CPU only:

* restructured (params.arithmetic = 1) vs reference-order Lagrange basis, field evaluation and push on 1e5 random inputs;
* the reference-order device functions against the oracle BIT FOR BIT — Lagrange basis incl. node hits, GetPositionInRefElem
  (Newton with start guesses 1, 3, 4) on deformed elements, ParticleInsideQuad3D — i.e. the claim "arithmetic = 0 is bitwise the
  oracle" (Arithmetic contract) checked on the code the kernels are built from.

* the device's TriaTracking (tria_hop of kernels.cuh: determinant tests, neighbour walk, periodic shift, specular reflection;
  tests/device_track_host.cpp) on the reference's own tracking checks: bitwise equal to the oracle at every step and in agreement
  with the state files the reference wrote (NIG_tracking_DSMC/periodic and ANSA_box, tests/test_reference_tracking.py);
* the device's RefMapping tracking (ref_tracking of ref.cuh: Newton in the old element, BC-side intersections on planar and
  bilinear sides, periodic shift, reflection, FIBGM relocation - with the stored candidate list and, for the reference's own
  background mesh with 125 elements in one cell, by repeated selection (select.cuh) - and the localisation fallback) on the same
  checks: positions, velocities, PartPosRef and elements bitwise equal to the oracle at every step.

The -m gpu parity tests remain the check of the kernels themselves."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
from oracle_lib import Oracle
from piclas_b200 import basis, hostmesh as hm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "device_math_host.cpp")
FLAGS = ["-std=c++17", "-O1", "-ffp-contract=off", "-I", os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include"),
         "-I", os.path.join(ROOT, "piclas_b200", "csrc")]
F64P, I32P = C.POINTER(C.c_double), C.POINTER(C.c_int32)


def _p(a, t=F64P):
    return a.ctypes.data_as(t)


def test_restructured_device_arithmetic_agrees_with_reference_order(tmp_path):
    exe = str(tmp_path / "device_math_host")
    subprocess.run(["g++"] + FLAGS + ["-o", exe, SRC], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("lagrange")


@pytest.fixture(scope="module")
def devmath(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("devmath") / "libdevmath.so")
    subprocess.run(["g++"] + FLAGS + ["-fPIC", "-shared", "-DDEVMATH_SHARED", "-o", so, SRC], check=True)
    return C.CDLL(so)


def test_device_lagrange_basis_is_the_oracles_bit_for_bit(devmath):
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (1, 1, 1), 3)
    orc = Oracle(mesh, cases.electron_params())
    rng = np.random.default_rng(0)
    for N in (1, 2, 3, 5, 7):
        xGP, wGP = basis.legendre_gauss_nodes_weights(N)
        wB = basis.barycentric_weights(xGP)
        pts = np.concatenate([rng.uniform(-1.2, 1.2, 300), xGP, xGP * (1 + 2e-16), [0.0, 1.0, -1.0]])
        for x in pts:
            L = np.zeros(N + 1)
            devmath.dm_lagrange(C.c_int(N + 1), C.c_double(x), _p(np.ascontiguousarray(xGP)), _p(np.ascontiguousarray(wB)), _p(L))
            assert np.array_equal(L, orc.lagrange(x, xGP, wB)), (N, x)
    orc.close()


@pytest.mark.parametrize("guess", [1, 3, 4])
def test_device_newton_and_inside_test_are_the_oracles_bit_for_bit(devmath, guess):
    lo, hi = [0, 0, 0], [1, 1, 1]
    mesh = hm.box_mesh(lo, hi, (3, 3, 2), 2, deform=cases.wavy(0.06, lo, hi))
    prm = cases.electron_params(RefMappingGuess=guess)
    orc = Oracle(mesh, prm)
    devmath.dm_set_newton_consts(_p(np.ascontiguousarray(mesh.XiCL_NGeo)), _p(np.ascontiguousarray(mesh.wBaryCL_NGeo)),
                                 C.c_double(prm.RefMappingEps), C.c_int(guess))
    rng = np.random.default_rng(guess)
    nchk = 0
    for e in range(mesh.nElems):
        first = mesh.ElemInfo[e, 4]                                    # ELEM_FIRSTNODEIND
        corners = np.ascontiguousarray(mesh.NodeCoords[first:first + 8])
        c0, c1 = corners.min(axis=0), corners.max(axis=0)
        x = np.ascontiguousarray(c0 + (c1 - c0) * rng.uniform(-0.15, 1.15, (400, 3)))       # inside, near and outside the element
        x[:8] = corners                                                                      # exactly on the corners
        x[8:14] = corners[mesh.ElemSideNodeID[e, :, :] - first].mean(axis=1)                 # side centres
        n = len(x)
        elem = np.full(n, e + 1, dtype=np.int32)
        # GetPositionInRefElem: ForceMode and isSuccessful as the deposition / interpolation call it
        xi_d, st_d = np.zeros((n, 3)), np.zeros(n, dtype=np.int32)
        devmath.dm_position_in_ref_elem(_p(np.ascontiguousarray(mesh.XCL_NGeo[e])), _p(np.ascontiguousarray(mesh.dXCL_NGeo[e])),
                                        _p(np.ascontiguousarray(mesh.ElemBaryNGeo[e])), _p(np.ascontiguousarray(mesh.XiEtaZetaBasis[e])),
                                        _p(np.ascontiguousarray(mesh.slenXiEtaZetaBasis[e])), C.c_int64(n), _p(x), C.c_int(1),
                                        _p(xi_d), _p(st_d, I32P))
        xi_o, suc_o, _ = orc.position_in_ref_elem(x, elem, force=True)
        assert np.array_equal(xi_d, xi_o), "reference coordinates differ from the oracle (element %d)" % (e + 1)
        assert np.array_equal(st_d & 1, suc_o)
        # ParticleInsideQuad3D
        side_node = np.ascontiguousarray((mesh.ElemSideNodeID[e] - first).astype(np.int32))
        conc = int(sum(int(bool(mesh.ConcaveElemSide[e, s])) << s for s in range(6)))
        ins_d, mask = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.uint32)
        devmath.dm_inside_quad3d(_p(corners), _p(side_node, I32P), C.c_int(conc), C.c_int64(n), _p(x), _p(ins_d, I32P),
                                 mask.ctypes.data_as(C.POINTER(C.c_uint32)))
        ins_o, det = orc.inside(x, elem)
        assert np.array_equal(ins_d, ins_o), "inside decision differs from the oracle (element %d)" % (e + 1)
        want = sum(((det[:, s, t] <= 0).astype(np.uint32) << np.uint32(2 * s + t)) for s in range(6) for t in range(2))
        assert np.array_equal(mask, want), "determinant signs differ from the oracle"
        nchk += n
    orc.close()
    assert nchk >= 7000


@pytest.fixture(scope="module")
def devtrack(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("devtrack") / "libdevtrack.so")
    subprocess.run(["g++"] + FLAGS + ["-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tests", "device_track_host.cpp")], check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("fast", [0, 1], ids=["reference-order", "restructured"])
@pytest.mark.parametrize("name", ["periodic", "ansa"])
def test_device_tria_tracking_is_the_oracles_and_reproduces_the_references_state_files(devtrack, name, fast):
    """fast = 1: the restructured arithmetic of params.arithmetic = 1 (inside test through the triangle planes with the exact
    fallback, exit-side shortcut on planar convex elements); it takes the same decisions, so the results stay bitwise equal."""
    import test_reference_tracking as trt
    mesh, prm, PD0, elem0, PD1, elem1, dt, nsteps = (trt.periodic_case if name == "periodic" else trt.ansa_case)(hm.TRIATRACKING)
    n = len(PD0)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    x, v, elem = np.ascontiguousarray(PD0[:, :3]), np.ascontiguousarray(PD0[:, 3:6]), i32(elem0)
    EI, SI, NC = i32(mesh.ElemInfo), i32(mesh.SideInfo), np.ascontiguousarray(mesh.NodeCoords)
    ESN, CC, bk, ba = i32(mesh.ElemSideNodeID), i32(mesh.ConcaveElemSide), i32(mesh.bc_kind), i32(mesh.bc_alpha)
    PV = np.ascontiguousarray(mesh.PeriodicVectors if mesh.nPeriodicVectors else np.zeros((1, 3)))
    orc = Oracle(mesh, prm)
    PSo, elo, spec = np.ascontiguousarray(PD0[:, :6]), elem0.copy(), PD0[:, 6].astype(np.int32)
    inside, isnew, E = np.ones(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros((mesh.nElems, 2, 2, 2, 3))
    status, most = np.zeros(n, dtype=np.int32), 0
    for it in range(nsteps):
        lp = x.copy()
        x = np.ascontiguousarray(x + v * dt)                                    # the push of a neutral particle (Leapfrog, q = 0)
        hops = devtrack.dt_tria_track(mesh.nElems, EI.shape[1], SI.shape[1], _p(EI, I32P), _p(SI, I32P), _p(NC), _p(ESN, I32P),
                                      _p(CC, I32P), mesh.nBCs, _p(bk, I32P), _p(ba, I32P), mesh.nPeriodicVectors, _p(PV), C.c_int64(n),
                                      _p(x), _p(lp), _p(v), _p(elem, I32P), _p(status, I32P), C.c_int(fast))
        assert hops >= 0 and not status.any(), (it, np.unique(status))
        most = max(most, hops)
        orc.push_track(dt, PSo, spec, elo, inside, isnew, E)
        assert np.array_equal(x, PSo[:, :3]) and np.array_equal(v, PSo[:, 3:]) and np.array_equal(elem, elo), "step %d" % it
    orc.close()
    assert most >= 3                                                             # multi-element flights were walked
    PS = np.concatenate([x, v], axis=1)
    if name == "periodic":
        trt.check_periodic(PS, elem, PD1, elem1, mesh.nElems)
    else:
        trt.check_ansa(PS, elem, PD0, PD1, elem1, mesh.nElems)


@pytest.mark.parametrize("name", ["periodic", "periodic-one-fibgm-cell", "ansa"])
def test_device_ref_tracking_is_the_oracles_and_reproduces_the_references_state_files(devtrack, name):
    import test_reference_tracking as trt
    from piclas_b200.abi import Marshalled
    if name == "ansa":
        case = trt.ansa_case(hm.REFMAPPING)
    else:
        case = trt.periodic_case(hm.REFMAPPING, **({"fibgm_deltas": (2.0, 1.0, 1.0)} if name.endswith("cell") else {}))
    mesh, prm, PD0, elem0, PD1, elem1, dt, nsteps = case
    if name.endswith("cell"):
        assert mesh.extra["FIBGM"]["nElems"].max() == 125                        # > REF_MAX_BGM: the repeated-selection path
    mar = Marshalled(mesh, prm)                                                   # the structs piclas_gpu_init receives
    n = len(PD0)
    x, v, elem = np.ascontiguousarray(PD0[:, :3]), np.ascontiguousarray(PD0[:, 3:6]), np.ascontiguousarray(elem0, dtype=np.int32)
    orc = Oracle(mesh, prm)
    PSo, elo, spec = np.ascontiguousarray(PD0[:, :6]), elem0.copy(), PD0[:, 6].astype(np.int32)
    xio, _, bad = orc.position_in_ref_elem(PSo[:, :3], elo)
    assert bad == 0
    xi = xio.copy()
    inside, isnew, E = np.ones(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros((mesh.nElems, 2, 2, 2, 3))
    status, relocated = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
    for it in range(nsteps):
        lp = x.copy()
        x = np.ascontiguousarray(x + v * dt)
        worst = devtrack.dt_ref_track(C.byref(mar.mesh), C.byref(mar.params), C.c_int64(n), _p(x), _p(lp), _p(v), _p(xi), _p(elem, I32P),
                                      _p(status, I32P), _p(relocated, I32P))
        assert worst == 0, (it, np.unique(status))
        orc.push_track(dt, PSo, spec, elo, inside, isnew, E, PartPosRef=xio)
        assert np.array_equal(x, PSo[:, :3]) and np.array_equal(v, PSo[:, 3:]) and np.array_equal(elem, elo), "step %d" % it
        assert np.array_equal(xi, xio), "PartPosRef differs from the oracle (step %d)" % it
    orc.close()
    PS = np.concatenate([x, v], axis=1)
    if name == "ansa":
        trt.check_ansa(PS, elem, PD0, PD1, elem1, mesh.nElems)
    else:
        trt.check_periodic(PS, elem, PD1, elem1, mesh.nElems)


@pytest.mark.parametrize("fast", [0, 1], ids=["reference-order", "restructured"])
@pytest.mark.parametrize("tag", ["wavy-periodic", "wavy-reflective", "cartesian-open-x"])
def test_device_tria_tracking_on_synthetic_meshes(devtrack, tag, fast):
    """Deformed elements (non-planar sides, concave flags), specular walls and open boundaries with flights of up to several
    elements: the device's crossing code on the host keeps particles, elements, positions and velocities bitwise equal to the oracle."""
    from piclas_b200.abi import TIMEDISC_LEAPFROG
    lo, hi = [0, 0, 0], [1, 1, 1]
    mesh = {"wavy-periodic": lambda: hm.box_mesh(lo, hi, (6, 5, 4), 1, deform=cases.wavy_periodic(0.04, lo, hi)),
            "wavy-reflective": lambda: hm.box_mesh(lo, hi, (5, 4, 4), 1, periodic=(False, False, False), wall_kind=hm.BC_REFLECTIVE,
                                                   deform=cases.wavy(0.05, lo, hi)),
            "cartesian-open-x": lambda: hm.box_mesh(lo, hi, (6, 6, 6), 1, periodic=(False, True, True), wall_kind=hm.BC_OPEN)}[tag]()
    prm = cases.electron_params(TimeDiscMethod=TIMEDISC_LEAPFROG, ChargeIC=(0.0,), DoInterpolation=0, DoDeposition=0)
    dt = 1e-8
    PS0, spec = cases.uniform_plasma(mesh, 20000, seed=5, vth_cells=0.7, dt=dt)
    orc = Oracle(mesh, prm)
    elem0 = orc.locate(PS0[:, :3]).astype(np.int32)
    keep = elem0 > 0
    PS0, spec, elem0 = np.ascontiguousarray(PS0[keep]), spec[keep], elem0[keep]
    n = len(spec)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    EI, SI, NC = i32(mesh.ElemInfo), i32(mesh.SideInfo), np.ascontiguousarray(mesh.NodeCoords)
    ESN, CC, bk, ba = i32(mesh.ElemSideNodeID), i32(mesh.ConcaveElemSide), i32(mesh.bc_kind), i32(mesh.bc_alpha)
    PV = np.ascontiguousarray(mesh.PeriodicVectors if mesh.nPeriodicVectors else np.zeros((1, 3)))
    x, v, elem = np.ascontiguousarray(PS0[:, :3]), np.ascontiguousarray(PS0[:, 3:]), elem0.copy()
    PSo, elo = PS0.copy(), elem0.copy()
    inside, isnew, E = np.ones(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros((mesh.nElems, 2, 2, 2, 3))
    alive = np.ones(n, dtype=bool)
    for it in range(6):
        lp = x.copy()
        x = np.ascontiguousarray(x + v * dt)
        idx = np.nonzero(alive)[0]
        xa, la, va = np.ascontiguousarray(x[idx]), np.ascontiguousarray(lp[idx]), np.ascontiguousarray(v[idx])
        ea, sa = np.ascontiguousarray(elem[idx]), np.zeros(len(idx), dtype=np.int32)
        hops = devtrack.dt_tria_track(mesh.nElems, EI.shape[1], SI.shape[1], _p(EI, I32P), _p(SI, I32P), _p(NC), _p(ESN, I32P),
                                      _p(CC, I32P), mesh.nBCs, _p(bk, I32P), _p(ba, I32P), mesh.nPeriodicVectors, _p(PV),
                                      C.c_int64(len(idx)), _p(xa), _p(la), _p(va), _p(ea, I32P), _p(sa, I32P), C.c_int(fast))
        assert hops >= 0 and set(np.unique(sa)) <= {0, 2}, np.unique(sa)           # TRK_OK or TRK_REMOVED (open boundary)
        x[idx], v[idx], elem[idx] = xa, va, ea
        alive[idx] = sa == 0
        orc.push_track(dt, PSo, spec, elo, inside, isnew, E)
        live = inside.astype(bool)
        assert np.array_equal(alive, live), "step %d: the set of removed particles differs" % it
        assert np.array_equal(elem[alive], elo[live]) and np.array_equal(x[alive], PSo[live, :3]) and np.array_equal(v[alive], PSo[live, 3:])
    orc.close()
    assert tag != "cartesian-open-x" or alive.sum() < n


@pytest.mark.parametrize("tag", ["wavy", "twisted-two-elements"])
def test_device_cvwm_particle_deposit_matches_the_oracle(devtrack, tag):
    """deposit_particle_general of kernels.cuh on the host: reference positions bitwise the oracle's, the same particles take the
    inverse-distance fallback, and the node sums built from the per-element accumulators agree with DepositionMethod_CVWM's raw
    NodeSource to round-off (the device sums per element first, the reference per node: different association)."""
    from piclas_b200.abi import Marshalled
    if tag == "wavy":
        lo, hi = [0, 0, 0], [1, 1, 1]
        mesh = hm.box_mesh(lo, hi, (4, 3, 3), 2, deform=cases.wavy_periodic(0.05, lo, hi))
        prm = cases.electron_params(MacroParticleFactor=(1e9,))
        PS, spec = cases.uniform_plasma(mesh, 6000, seed=12, vth_cells=0.3, dt=1e-8)
        orc = Oracle(mesh, prm)
        elem = orc.locate(PS[:, :3]).astype(np.int32)
        keep = elem > 0
        PS, spec, elem = np.ascontiguousarray(PS[keep]), spec[keep], elem[keep]
    else:
        mesh, prm, PS, spec = cases.plasma_ball_two_elements(True)
        orc = Oracle(mesh, prm)
        elem = np.ascontiguousarray(orc.locate(PS[:, :3]), dtype=np.int32)
        assert (elem > 0).all()
    n = len(spec)
    mar = Marshalled(mesh, prm)
    acc, xi, failed = np.zeros((mesh.nElems, 8, 4)), np.zeros((n, 3)), np.zeros(n, dtype=np.int32)
    assert devtrack.dt_cvwm_accumulate(C.byref(mar.mesh), C.byref(mar.params), C.c_int64(n), _p(np.ascontiguousarray(PS)),
                                       _p(np.ascontiguousarray(spec, dtype=np.int32), I32P), _p(elem, I32P), _p(acc), _p(xi),
                                       _p(failed, I32P)) == 0
    xi_o, suc_o, _ = orc.position_in_ref_elem(PS[:, :3], elem, force=True)
    assert np.array_equal(xi, xi_o) and np.array_equal(failed, 1 - suc_o)
    if tag != "wavy":
        assert failed.sum() > 100                                       # the twisted interface: SucRefPos = F for many particles
    NS = np.zeros((mesh.nUniqueNodes, 4))
    node = mesh.NodeInfo[mesh.ElemNodeID - 1] - 1                       # [nElems][8] unique node of CGNS corner n
    np.add.at(NS, node.reshape(-1), acc.reshape(-1, 4))
    if mesh.nPeriodicVectors == 0:
        ref = orc.deposit_raw(PS, spec, elem, np.ones(n, dtype=np.int32))
        for c in range(4):
            scale = np.abs(ref[:, c]).max()
            if scale > 0:
                assert np.abs(NS[:, c] - ref[:, c]).max() <= 1e-13 * scale, c
    q = prm.ChargeIC[0] * prm.MacroParticleFactor[0]
    assert abs(NS[:, 3].sum() - n * q) <= 1e-12 * abs(n * q)          # the weights of every particle sum to one
    orc.close()


def _hint_vs_exact(devtrack, mesh, x, lp, elem):
    """far_hint_record (csrc/hint.cuh) and the device's exact walk (reference order), both on the host, for the flights lp -> x."""
    n = len(x)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    EI, SI, NC = i32(mesh.ElemInfo), i32(mesh.SideInfo), np.ascontiguousarray(mesh.NodeCoords)
    ESN, CC, bk, ba = i32(mesh.ElemSideNodeID), i32(mesh.ConcaveElemSide), i32(mesh.bc_kind), i32(mesh.bc_alpha)
    PV = np.ascontiguousarray(mesh.PeriodicVectors if mesh.nPeriodicVectors else np.zeros((1, 3)))
    args = (mesh.nElems, EI.shape[1], SI.shape[1], _p(EI, I32P), _p(SI, I32P), _p(NC), _p(ESN, I32P), _p(CC, I32P), mesh.nBCs,
            _p(bk, I32P), _p(ba, I32P), mesh.nPeriodicVectors, _p(PV), C.c_int64(n))
    xe, lpe, ve, ele, st = x.copy(), lp.copy(), np.zeros_like(x), i32(elem).copy(), np.zeros(n, dtype=np.int32)
    hops = devtrack.dt_tria_track(*args, _p(xe), _p(lpe), _p(ve), _p(ele, I32P), _p(st, I32P), C.c_int(0))
    assert hops >= 0
    xh, lph, fin = x.copy(), lp.copy(), np.zeros(n, dtype=np.int32)
    devtrack.dt_tria_hint.restype = C.c_int64
    settled = devtrack.dt_tria_hint(*args, _p(xh), _p(lph), _p(i32(elem), I32P), _p(fin, I32P))
    assert settled == int((fin > 0).sum())
    s = fin > 0
    assert not (st[s] != 0).any(), "the hint settled a particle the exact walk removes or loses"
    assert np.array_equal(fin[s], ele[s]), "the hint's element differs from the exact walk's"
    L = np.abs(mesh.NodeCoords).max()
    if s.any():
        assert np.abs(xh[s] - xe[s]).max() <= 1e-12 * L                  # periodic displacement: x + vector vs crossing point + rest
    return fin, ele, st


@pytest.mark.parametrize("tag", ["periodic-box", "open-x", "thin-periodic", "wavy"])
def test_far_hint_decisions_are_the_exact_walks(devtrack, tag):
    """k_far_hint's per-record decision, compiled for the host, on flights of up to several elements: whenever it settles a
    record, the element (and the periodically displaced position) is what SingleParticleTriaTracking3D finds; what it cannot
    decide with its margins it leaves alone.  Includes flights that graze edges and corners and end points on element faces."""
    rng = np.random.default_rng(20261018)
    lo, hi = np.zeros(3), np.ones(3)
    if tag == "periodic-box":
        mesh = hm.box_mesh(lo, hi, (6, 5, 4), 1)
    elif tag == "open-x":
        mesh = hm.box_mesh(lo, hi, (6, 5, 4), 1, periodic=(False, True, True))
    elif tag == "thin-periodic":
        mesh = hm.box_mesh(lo, hi, (8, 1, 2), 1)                           # an element is its own neighbour in y
    else:
        mesh = hm.box_mesh(lo, hi, (5, 4, 4), 1, deform=cases.wavy(0.05, lo, hi))   # non-planar inner sides: nothing to decide
    ne = np.array(mesh.extra["nelems"])
    h = 1.0 / ne
    n = 60000
    lp = rng.random((n, 3))
    step = rng.normal(0.0, 0.6, (n, 3)) * h                                # up to three crossings are common
    # adversarial thirds: start points near element corners (flights through edge / corner regions), end points exactly on faces
    k = n // 3
    corner = np.round(lp[:k] / h) * h
    lp[:k] = np.clip(corner + rng.normal(0.0, 0.02, (k, 3)) * h, 1e-9, 1 - 1e-9)
    x = lp + step
    x[k:2 * k, 0] = np.round(x[k:2 * k, 0] / h[0]) * h[0]
    if tag == "wavy":
        orc = Oracle(mesh, cases.electron_params())
        elem = orc.locate(lp)
        orc.close()
    else:
        elem = hm.cartesian_locate(mesh, lp)
    fin, ele, st = _hint_vs_exact(devtrack, mesh, np.ascontiguousarray(x), np.ascontiguousarray(lp), elem)
    share = (fin > 0).mean()
    if tag == "wavy":
        assert share < 0.2                                                 # planar elements exist only where the deformation vanishes
    else:
        assert share > {"periodic-box": 0.5, "open-x": 0.35, "thin-periodic": 0.3}[tag], share   # up to three crossings are decided by the planes
        left = (fin == 0) & (st == 0)
        assert left.any()                                                  # ... and the margins did leave records to the exact walk


@pytest.mark.parametrize("tracking", ["triatracking", "refmapping"])
@pytest.mark.parametrize("kind", ["sin_deviation", "cos_distribution"])
def test_device_emission_on_the_host(devtrack, tracking, kind):
    """csrc/emit.cuh compiled for the host: the lattice positions are bitwise the harness restatements' (the sin_deviation one is
    pinned on the reference's restart file, tests/test_reference_emission.py) and single_point_to_element finds the element
    SinglePointToElement does (FIBGM cell, radius filter, nearest barycentre first; lattice points on faces included), for the
    whole mesh and for the element range of one rank."""
    from piclas_b200.abi import Marshalled, DEPO_SF
    lo, hi = [0.0, 0.0, 0.0], [6.2831, 0.9, 0.4]
    ref = tracking == "refmapping"
    mesh = hm.box_mesh(lo, hi, (6, 3, 2), 2, tracking=hm.REFMAPPING if ref else hm.TRIATRACKING,
                       deform=cases.wavy(0.04, lo, hi) if ref else None)
    hm.add_fibgm(mesh)
    if ref:
        hm.add_refmapping_tables(mesh)
        prm = cases.electron_params(TrackingMethod=hm.REFMAPPING, DepositionType=DEPO_SF, DoDeposition=0)
    else:
        prm = cases.electron_params()
    n3, amp, wn = ((25, 6, 4), 0.01, 2.0) if kind == "sin_deviation" else ((40, 3, 3), 0.05, 0.5)
    want = (cases.sin_deviation if kind == "sin_deviation" else cases.cos_distribution)(mesh.xyz_min, mesh.xyz_max, *n3, amp, wn)
    mar = Marshalled(mesh, prm)
    orc = Oracle(mesh, prm)
    n = len(want)
    for first, last in ((1, mesh.nElems), (13, 24)):
        X, el = np.zeros((n, 3)), np.zeros(n, dtype=np.int32)
        rc = devtrack.dt_emit_lattice(C.byref(mar.mesh), C.byref(mar.params), 1 if kind == "sin_deviation" else 2, *n3, C.c_double(amp),
                                      C.c_double(wn), first, last, _p(X), _p(el, I32P))
        assert rc == 0
        assert np.array_equal(X, want)
        assert np.array_equal(el, cases.single_point_to_element(mesh, orc, want, first=first, last=last, refmapping=ref))
        assert (el > 0).sum() >= n // 4
    orc.close()
