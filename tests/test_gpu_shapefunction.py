"""GPU parity + known answers for the shape-function depositions."""
import json
import os

import numpy as np
import pytest

import cases
from oracle_lib import Oracle
from piclas_b200 import hostmesh as hm
from piclas_b200.abi import Params, DEPO_SF, DEPO_SF_CC, DEPO_SF_ADAPTIVE

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_known_answers.json")))["NIG_PIC_Deposition/Plasma_Ball_Shape-function"]
RTOL = 1e-12


def _deposit_both(mesh, prm, PS, spec, elem):
    from piclas_b200.particle_step import ParticleStep
    orc = Oracle(mesh, prm)
    PSo, _ = orc.deposit(PS, spec, elem, np.ones(len(spec), dtype=np.int32))
    with ParticleStep(mesh, prm) as gpu:
        gpu.UploadParticles(PS, spec, elem)
        PSg, _ = gpu.Deposition()
    q = orc.deposited_charge(PSg)
    orc.close()
    for c in range(4):
        scale = np.abs(PSo[..., c]).max()
        if scale > 0:
            assert np.abs(PSg[..., c] - PSo[..., c]).max() / scale <= RTOL, c
        else:
            assert not PSg[..., c].any()
    return q


def plasma_ball(nelems=(4, 4, 1), n=100000):
    mesh = hm.box_mesh([-5, -5, -5], [5, 5, 5], tuple(nelems), 3)
    hm.add_fibgm(mesh)
    rng = np.random.default_rng(20261017)
    x = cases.sphere_points(rng, n, 0.5)
    PS = np.zeros((n, 6))
    PS[:, :3] = x
    PS[:, 3:] = rng.normal(0, 1.0, (n, 3))
    return mesh, PS, np.ones(n, dtype=np.int32), hm.cartesian_locate(mesh, x)


@pytest.mark.parametrize("kind", ["sf", "cc", "adaptive"])
@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["case"].replace(" ", "-") for c in GOLD["cases"]])
def test_plasma_ball_shape_function_known_answers(case, kind):
    """NIG_PIC_Deposition/Plasma_Ball_Shape-function-{x,y,z}Dir: periodic box of 16 elements, N=3, r_sf=2, alpha=2, 100000
    particles; all 18 reference values (1-D / 2-D in every direction; shape_function 5 % relative, _cc and _adaptive 1e-9)."""
    mesh, PS, spec, elem = plasma_ball(case["nelems"])
    prm = Params(ChargeIC=(1.60217653e-5,), MassIC=(1.0,), MacroParticleFactor=(200.0,),
                 DepositionType={"sf": DEPO_SF, "cc": DEPO_SF_CC, "adaptive": DEPO_SF_ADAPTIVE}[kind])
    if kind == "adaptive":
        hm.shape_function_adaptive_setup(mesh, prm, 2, dim_sf=case["dim"], dim_sf_dir=case["dir"], sfDepo3D=True, smoothing=True)
    else:
        hm.shape_function_setup(mesh, prm, 2.0, 2, dim_sf=case["dim"], dim_sf_dir=case["dir"], sfDepo3D=True)
    q = _deposit_both(mesh, prm, PS, spec, elem)
    tol = GOLD["tolerances"][kind]
    assert abs(q - case[kind]) <= (tol["value"] * case[kind] if tol["type"] == "relative" else tol["value"])


@pytest.mark.parametrize("kind", [DEPO_SF, DEPO_SF_CC])
@pytest.mark.parametrize("periodic", [True, False])
def test_shape_function_3d(kind, periodic):
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (5, 4, 3), 2, periodic=(periodic,) * 3)
    hm.add_fibgm(mesh)
    prm = cases.electron_params(DepositionType=kind, MacroParticleFactor=(1e9,))
    hm.shape_function_setup(mesh, prm, 0.27, 3, dim_sf=3)
    PS, spec = cases.uniform_plasma(mesh, 6000, seed=17)
    elem = hm.cartesian_locate(mesh, PS[:, :3])
    q = _deposit_both(mesh, prm, PS, spec, elem)
    if kind == DEPO_SF_CC:
        assert abs(q - 6000 * 1e9 * (-cases.QE)) <= 1e-9 * abs(6000 * 1e9 * cases.QE)


def test_shape_function_adaptive_radius_per_element():
    """shape_function_adaptive (periodic / smoothing path): the radius is the one of the particle's element."""
    mesh = hm.box_mesh([0, 0, 0], [3, 1, 1], (12, 2, 2), 3)
    hm.add_fibgm(mesh)
    rng = np.random.default_rng(4)
    r = 0.3 + 0.1 * rng.random(mesh.nElems)
    mesh.extra["SFElemr2"] = np.ascontiguousarray(np.stack([r, r * r], axis=1))
    prm = cases.electron_params(DepositionType=DEPO_SF_ADAPTIVE, MacroParticleFactor=(1e9,))
    hm.shape_function_setup(mesh, prm, 1.0, 4, dim_sf=1, dim_sf_dir=1, sfDepo3D=False)
    PS, spec = cases.uniform_plasma(mesh, 5000, seed=23)
    elem = hm.cartesian_locate(mesh, PS[:, :3])
    q = _deposit_both(mesh, prm, PS, spec, elem)
    # line deposition (3D-deposition = F): the integral over the volume is charge * dimFactorSF
    assert abs(q - 5000 * 1e9 * (-cases.QE) * prm.dimFactorSF) <= 1e-9 * abs(5000 * 1e9 * cases.QE)


def test_push_after_shape_function_deposit_uses_newton_in_push_kernel():
    """After a shape-function deposit no reference positions are cached: the push kernel maps x -> xi itself."""
    from piclas_b200.particle_step import ParticleStep
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (4, 4, 4), 2)
    hm.add_fibgm(mesh)
    prm = cases.electron_params(DepositionType=DEPO_SF)
    hm.shape_function_setup(mesh, prm, 0.2, 2, dim_sf=3)
    dt = 1e-8
    PS, spec = cases.uniform_plasma(mesh, 5000, seed=3, vth_cells=0.3, dt=dt)
    elem = hm.cartesian_locate(mesh, PS[:, :3])
    E = cases.smooth_field(mesh, 1e-4)
    n = len(spec)
    orc = Oracle(mesh, prm)
    PSo, elo = PS.copy(), elem.copy()
    inside, isnew = np.ones(n, dtype=np.int32), np.ones(n, dtype=np.int32)
    with ParticleStep(mesh, prm) as gpu:
        gpu.UploadParticles(PS, spec, elem, IsNewPart=isnew, ids=np.arange(n))
        gpu.SetField(E)
        for _ in range(3):
            gpu.Deposition()
            gpu.PushAndTrack(dt)
            orc.push_track(dt, PSo, spec, elo, inside, isnew, E)
        d = gpu.DownloadParticles()
    o = np.argsort(d["ids"])
    assert np.array_equal(d["GlobalElemID"][o], elo)
    assert np.abs(d["PartState"][o] - PSo).max() / np.abs(PSo).max() <= RTOL
