"""bench.py's full-size property checks (charge conservation, particle counts) on a small case, CPU only: the helper restates
CalcDepositedCharge (pic_analyze.f90:165-175) and must agree with the oracle's."""
import numpy as np

import bench
import cases
from oracle_lib import Oracle
from piclas_b200 import hostmesh as hm


def test_deposited_charge_helper_and_check_record():
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (6, 5, 4), 3)
    prm = cases.electron_params()
    PS, spec = cases.uniform_plasma(mesh, 20000, seed=3, vth_cells=0.3, dt=1e-8)
    elem = hm.cartesian_locate(mesh, PS[:, :3])
    orc = Oracle(mesh, prm)
    src, _ = orc.deposit(PS, spec, elem, np.ones(len(spec), dtype=np.int32))
    q = len(spec) * prm.ChargeIC[0] * prm.MacroParticleFactor[0]
    assert abs(bench.deposited_charge(mesh, src[..., 3]) - orc.deposited_charge(src)) <= 1e-14 * abs(q)
    assert abs(bench.deposited_charge(mesh, src[..., 3]) - q) <= 1e-13 * abs(q)          # cell_volweight_mean conserves charge

    class Step:                                                                           # stands in for ParticleStep
        def Deposition(self, **kw):
            pass

        def ChargeDensity(self, out):
            out[...] = src[..., 3]

        def NumParticles(self):
            return len(spec)

        def KineticEnergy(self):
            return np.array([1.0]), np.array([len(spec)])

    rec = bench.full_size_checks(Step(), mesh, len(spec), prm.ChargeIC[0] * prm.MacroParticleFactor[0])
    assert rec["ok"] and rec["charge_conservation_rel_err"] <= 1e-13 and rec["particles_in_reduction"] == len(spec)
    assert bench.full_size_checks(Step(), mesh, len(spec) + 1, 1.0)["ok"] is False        # a lost particle is reported
    assert bench.full_size_checks(object(), mesh, 1, 1.0)["ok"] is False                  # and a failing check never raises
