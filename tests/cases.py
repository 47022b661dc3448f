"""Seeded synthetic cases shared by the CPU and GPU tests (shapes follow BASELINE.json `configs`)."""
from __future__ import annotations

import numpy as np

from piclas_b200 import hostmesh as hm
from piclas_b200.abi import Params, DEPO_CVWM, TIMEDISC_BORIS_LEAPFROG, TIMEDISC_LEAPFROG

QE = 1.60217653e-19
ME = 9.1093826e-31


def sphere_points(rng, n, r):
    pts = np.zeros((0, 3))
    while pts.shape[0] < n:
        q = rng.uniform(-r, r, (2 * n, 3))
        q = q[(q * q).sum(axis=1) <= r * r]
        pts = np.concatenate([pts, q])
    return pts[:n].copy()


def plasma_ball_cvwm():
    """regressioncheck/NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean: 10^3 box [-1,1]^3, N=1, TriaTracking, CVWM,
    3333 particles in a sphere r=0.5, q=1.60217653e-5, MPF=200 (parameter.ini, hopr.ini)."""
    mesh = hm.box_mesh([-1, -1, -1], [1, 1, 1], (10, 10, 10), 1)
    prm = Params(ChargeIC=(1.60217653e-5, -QE), MassIC=(1.0, ME), MacroParticleFactor=(200.0, 200.0),
                 DepositionType=DEPO_CVWM, carryParticleIDs=1)
    rng = np.random.default_rng(20261017)
    x = sphere_points(rng, 3333, 0.5)
    PS = np.zeros((3333, 6))
    PS[:, :3] = x
    return mesh, prm, PS, np.ones(3333, dtype=np.int32), hm.cartesian_locate(mesh, x)


def reference_plasma_ball():
    """The reference's own particles and its deposited charge density (tests/golden/make_reference_vectors.py): returns the
    10^3 mesh, parameters, PartState, species, element ids and DG_Source(4,:) reordered from HOPR's element order to ours."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plasma_ball_cvwm_reference.npz"))
    mesh = hm.box_mesh([-1, -1, -1], [1, 1, 1], (10, 10, 10), 1)
    prm = Params(ChargeIC=(1.60217653e-5, -QE), MassIC=(1.0, ME), MacroParticleFactor=(200.0, 200.0),
                 DepositionType=DEPO_CVWM, carryParticleIDs=1)
    ijk = np.floor((g["ElemBarycenters"] + 1.0) / 0.2).astype(np.int64)
    ours = ijk[:, 0] + 10 * (ijk[:, 1] + 10 * ijk[:, 2])            # HOPR element -> our element
    assert len(np.unique(ours)) == 1000
    rho = np.empty((1000, 2, 2, 2))
    rho[ours] = g["DG_Source_charge"]
    PS = np.ascontiguousarray(g["PartData"][:, :6])
    spec = g["PartData"][:, 6].astype(np.int32)
    return mesh, prm, PS, spec, hm.cartesian_locate(mesh, PS[:, :3]), rho, g


def plasma_ball_two_elements(deformed):
    """regressioncheck/NIG_PIC_Deposition/Plasma_Ball_cell_volweight_mean_save_CVWM: two hexahedra in [-1,1]^3 (corners of
    hopr.ini; `deformed`: the shared side is strongly twisted, which triggers the SucRefPos=F inverse-distance fallback of
    DepositionMethod_CVWM), reflective walls, N=1, TriaTracking, 3333 particles in a sphere r=0.5."""
    lx = 0.4 if deformed else 0.0
    li = 1.0
    z1 = [(-li, -li, -li), (lx, -li, -li), (-lx, li, -li), (-li, li, -li), (-li, -li, li), (-lx, -li, li), (lx, li, li), (-li, li, li)]
    z2 = [(lx, -li, -li), (li, -li, -li), (li, li, -li), (-lx, li, -li), (-lx, -li, li), (li, -li, li), (li, li, li), (lx, li, li)]
    pts, ids = [], {}
    elems = []
    for zone in (z1, z2):
        e = []
        for c in zone:
            if c not in ids:
                ids[c] = len(pts)
                pts.append(c)
            e.append(ids[c])
        elems.append(e)
    mesh = hm.build_mesh(np.array(pts, float), np.array(elems), 1, [(hm.BC_REFLECTIVE, 0)],
                         lambda cen, fid: np.ones(len(cen), dtype=np.int32))
    prm = Params(ChargeIC=(1.60217653e-5, -QE), MassIC=(1.0, ME), MacroParticleFactor=(200.0, 200.0),
                 DepositionType=DEPO_CVWM, carryParticleIDs=1)
    rng = np.random.default_rng(20261018)
    x = sphere_points(rng, 3333, 0.5)
    PS = np.zeros((3333, 6))
    PS[:, :3] = x
    return mesh, prm, PS, np.ones(3333, dtype=np.int32)


def smooth_field(mesh, amp=1.0):
    """Smooth analytic E sampled at the Gauss points, [nElems,k,j,i,3]."""
    X = mesh.Elem_xGP
    lo, hi = mesh.xyz_min, mesh.xyz_max
    L = hi - lo
    s = 2 * np.pi * (X - lo) / L
    E = np.zeros(X.shape)
    E[..., 0] = amp * np.sin(s[..., 0]) * np.cos(s[..., 1])
    E[..., 1] = amp * 0.5 * np.cos(s[..., 1] + 0.3) * np.sin(s[..., 2])
    E[..., 2] = amp * 0.25 * np.sin(s[..., 2] + s[..., 0])
    return np.ascontiguousarray(E)


def wavy(amplitude, lo, hi):
    """Smooth deformation that vanishes on the box boundary (keeps periodic faces congruent, faces become non-planar)."""
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)

    def f(c):
        s = np.pi * (c - lo) / (hi - lo)
        bump = np.sin(s[:, 0]) * np.sin(s[:, 1]) * np.sin(s[:, 2])
        out = c.copy()
        out[:, 0] += amplitude * (hi[0] - lo[0]) * bump * np.cos(3 * s[:, 1])
        out[:, 1] += amplitude * (hi[1] - lo[1]) * bump * np.cos(2 * s[:, 2] + 0.5)
        out[:, 2] += amplitude * (hi[2] - lo[2]) * bump * np.cos(2 * s[:, 0] + 1.0)
        return out
    return f


def wavy_periodic(amplitude, lo, hi):
    """Deformation that does NOT vanish on the boundary but is periodic: the x-faces become congruent bilinear patches, the y- and
    z-faces planar non-rectangular (RefMapping: ComputeBiLinearIntersection for BILINEAR and PLANAR_NONRECT BC sides)."""
    lo = np.asarray(lo, float)
    hi = np.asarray(hi, float)

    def f(c):
        s = 2.0 * np.pi * (c - lo) / (hi - lo)
        out = c.copy()
        out[:, 0] += amplitude * (hi[0] - lo[0]) * np.cos(s[:, 1]) * np.cos(s[:, 2])
        return out
    return f


def uniform_plasma(mesh, n, seed, vth_cells=0.2, dt=1.0, species=1, nspecies=1):
    """Uniform positions, Maxwellian velocities with sigma*dt = vth_cells * h (SURVEY.md §8d synthetic input)."""
    rng = np.random.default_rng(seed)
    lo, hi = mesh.xyz_min, mesh.xyz_max
    x = lo + (hi - lo) * rng.random((n, 3))
    h = ((hi - lo) / np.array(mesh.extra["nelems"])).min()
    v = rng.normal(0.0, vth_cells * h / dt, (n, 3))
    PS = np.ascontiguousarray(np.concatenate([x, v], axis=1))
    if nspecies > 1:
        spec = rng.integers(1, nspecies + 1, n).astype(np.int32)
    else:
        spec = np.full(n, species, dtype=np.int32)
    return PS, spec


def sin_deviation(xmin, xmax, nx, ny, nz, amplitude, wavenumber):
    """SetParticlePositionSinDeviation (particle_emission_tools.f90:1235-1293), loop order and expressions as written there:
    x = xmin + x_pos + A sin(k 2 pi / xlen * x_pos) with x_pos = i x_step - x_step / 2; y, z on the cell centres."""
    xmin, xmax = np.asarray(xmin, dtype=np.float64), np.asarray(xmax, dtype=np.float64)
    xlen, ylen, zlen = np.abs(xmax - xmin)
    pilen = 2.0 * np.pi / xlen
    x_step, y_step, z_step = xlen / nx, ylen / ny, zlen / nz
    pos = []
    for i in range(1, nx + 1):
        x_pos = i * x_step - x_step * 0.5
        x = xmin[0] + x_pos + amplitude * np.sin(wavenumber * pilen * x_pos)
        for j in range(1, ny + 1):
            y = xmin[1] + j * y_step - y_step * 0.5
            for k in range(1, nz + 1):
                pos.append((x, y, xmin[2] + k * z_step - z_step * 0.5))
    return np.array(pos)


def cos_distribution(xmin, xmax, nx, ny, nz, amplitude, wavenumber):
    """SetParticlePositionCosDistribution (particle_emission_tools.f90:1299-1371): x from the inverse of F(x) = x + a/w sin(w x)
    by Newton's method (residual <= 1e-12), y and z on the cell centres; loop order as written there."""
    xmin, xmax = np.asarray(xmin, dtype=np.float64), np.asarray(xmax, dtype=np.float64)
    xlen, ylen, zlen = np.abs(xmax - xmin)
    x_step, y_step, z_step = xlen / nx, ylen / ny, zlen / nz
    a, w = amplitude, wavenumber
    pos = []
    for i in range(1, nx + 1):
        x_uniform = i * x_step - x_step * 0.5
        x_pos = x_uniform
        while abs(x_pos + a / w * np.sin(w * x_pos) - x_uniform) > 1e-12:
            x_pos = x_pos - (x_pos + a / w * np.sin(w * x_pos) - x_uniform) / (1 + a * np.cos(w * x_pos))
        x = xmin[0] + x_pos
        for j in range(1, ny + 1):
            y = xmin[1] + j * y_step - y_step * 0.5
            for k in range(1, nz + 1):
                pos.append((x, y, xmin[2] + k * z_step - z_step * 0.5))
    return np.array(pos)


def single_point_to_element(mesh, orc, X, first=None, last=None, refmapping=False):
    """SinglePointToElement(doHALO=F) (particle_localization.f90:81-190) for many points: FIBGM cell by CEILING, the cell's
    elements within ElemRadius2NGeo of the point, nearest barycentre first (stable order), the first one that holds the point
    (ParticleInsideQuad3D, or MAXVAL|xi| <= ElemEpsOneCell under RefMapping) among the elements first..last; -1 otherwise.
    The element tests are the oracle's."""
    fb = mesh.extra["FIBGM"]
    first = 1 if first is None else first
    last = mesh.nElems if last is None else last
    X = np.ascontiguousarray(X, dtype=np.float64)
    out = np.full(len(X), -1, dtype=np.int32)
    ni, nj, nk = fb["max"]
    nE, off, El = fb["nElems"].reshape(-1), fb["offsetElem"].reshape(-1), fb["Element"]
    for p, x in enumerate(X):
        c = np.ceil((x - mesh.xyz_min) / fb["deltas"]).astype(np.int64)
        c = np.maximum(np.minimum(np.array([ni, nj, nk]), c), 1)
        cell = (c[0] - 1) + ni * ((c[1] - 1) + nj * (c[2] - 1))
        cand = El[off[cell]:off[cell] + nE[cell]]
        d2 = ((x[None, :] - mesh.ElemBaryNGeo[cand - 1]) ** 2).sum(axis=1)
        keep = d2 <= mesh.ElemRadius2NGeo[cand - 1]
        order = np.argsort(np.where(keep, d2, -1.0), kind="stable")
        for i in order:
            e = int(cand[i])
            if not keep[i] or e < first or e > last:
                continue
            if refmapping:
                xi, suc, _ = orc.position_in_ref_elem(x[None, :], np.array([e], dtype=np.int32), force=False)
                ok = bool(suc[0]) and np.abs(xi[0]).max() <= mesh.extra["ElemEpsOneCell"][e - 1]
            else:
                ok = bool(orc.inside(x[None, :], np.array([e], dtype=np.int32))[0][0])
            if ok:
                out[p] = e
                break
    return out


def electron_params(**kw):
    d = dict(ChargeIC=(-QE,), MassIC=(ME,), MacroParticleFactor=(1.0e3,), DepositionType=DEPO_CVWM,
             TimeDiscMethod=TIMEDISC_BORIS_LEAPFROG, carryParticleIDs=1)
    d.update(kw)
    return Params(**d)
