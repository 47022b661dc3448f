"""Multi-rank worker (run under torch.distributed.run): N ranks step an element-partitioned plasma, rank 0 checks the
result against the single-rank oracle — the reference's own criterion for its MPI=1,2,5,10 regression runs
(regressioncheck/*/command_line.ini): the multi-rank result must agree with the single-rank one.

  --engine gpu     ranks drive libpiclas_gpu.so through piclas_b200.multi.ParticleStepRank (NCCL)
  --engine oracle  ranks drive the CPU oracle; exercises the same transport functions of piclas_b200.multi over gloo
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle_lib import Oracle  # noqa: E402
from piclas_b200 import hostmesh as hm  # noqa: E402
from piclas_b200 import multi  # noqa: E402
from piclas_b200.abi import DEPO_CVWM, DEPO_SF, DEPO_SF_CC  # noqa: E402


class OracleRank:
    """CPU stand-in for one rank: host AoS arrays as in the Fortran module globals, oracle operators, same exchanges."""

    def __init__(self, mesh, prm, rank, world):
        self.rank, self.world = rank, world
        self.off = hm.partition(mesh, world)
        self.mesh = mesh
        self.orc = Oracle(mesh, prm, offsetElem=int(self.off[rank]), nElems=int(self.off[rank + 1] - self.off[rank]))
        self.dev = torch.device("cpu")
        self.sf = prm.DepositionType != DEPO_CVWM
        if self.sf:
            self.orc_full = Oracle(mesh, prm)

    def upload(self, PS, spec, elem, ids, isnew=1):
        self.PS, self.spec, self.elem, self.ids = PS.copy(), spec.copy(), elem.copy(), ids.copy()
        self.isnew = np.full(len(spec), isnew, dtype=np.int32)

    def set_field(self, E):
        self.E = np.ascontiguousarray(E[int(self.off[self.rank]):int(self.off[self.rank + 1])])

    def deposition(self):
        inside = np.ones(len(self.spec), dtype=np.int32)
        if self.sf:
            # the local particles' contributions to every element, then the DOF halo: the blocks of foreign elements go to
            # their owners, which add them rank after rank (pic_depo_method.f90:940-996)
            full, _ = self.orc_full.deposit(self.PS, self.spec, self.elem, inside)
            off, w = self.off, self.world
            dpe = int(np.prod(full.shape[1:]))
            others = [r for r in range(w) if r != self.rank]
            send = [int(off[r + 1] - off[r]) if r != self.rank else 0 for r in range(w)]
            mine = int(off[self.rank + 1] - off[self.rank])
            recv = [mine if r != self.rank else 0 for r in range(w)]
            sb = np.concatenate([full[int(off[r]):int(off[r + 1])].reshape(-1) for r in others]) if others else np.zeros(0)
            sbuf = torch.from_numpy(np.ascontiguousarray(sb))
            rbuf = torch.empty(sum(recv) * dpe, dtype=torch.float64)
            multi.exchange_sf_halo(sbuf, send, rbuf, recv, dpe)
            out = full[int(off[self.rank]):int(off[self.rank + 1])].copy()
            rb = rbuf.numpy().reshape((len(others),) + out.shape)
            for i in range(len(others)):
                out = out + rb[i]
            return out, None
        NS = self.orc.deposit_raw(self.PS, self.spec, self.elem, inside)
        t = torch.from_numpy(NS.reshape(-1))
        multi.halo_sum(t)
        return self.orc.deposit_finish(NS)

    def push_track(self, dt):
        n = len(self.spec)
        inside = np.ones(n, dtype=np.int32)
        nl, _, _ = self.orc.push_track(dt, self.PS, self.spec, self.elem, inside, self.isnew, self.E)
        rank_of = self.mesh.ElemInfo[:, 6]
        alive = inside.astype(bool)
        dest = np.where(alive, rank_of[np.maximum(self.elem, 1) - 1], -1)
        stay = alive & (dest == self.rank)
        send_counts, bufs = [], []
        for r in range(self.world):
            m = alive & (dest == r) & (r != self.rank)
            send_counts.append(int(m.sum()))
            b = np.zeros((int(m.sum()), 9))
            b[:, :6] = self.PS[m]
            b[:, 6] = self.spec[m]
            b[:, 7] = self.elem[m]
            b[:, 8] = self.ids[m].astype(np.int64).view(np.float64)
            bufs.append(b)
        sbuf = torch.from_numpy(np.ascontiguousarray(np.concatenate(bufs).reshape(-1)))
        recv_counts = multi.exchange_counts(send_counts, self.dev)
        rbuf = torch.empty(sum(recv_counts) * 9, dtype=torch.float64)
        multi.exchange_particles(sbuf, send_counts, rbuf, recv_counts, 9)
        rb = rbuf.numpy().reshape(-1, 9)
        self.PS = np.ascontiguousarray(np.concatenate([self.PS[stay], rb[:, :6]]))
        self.spec = np.concatenate([self.spec[stay], rb[:, 6].astype(np.int32)])
        self.elem = np.concatenate([self.elem[stay], rb[:, 7].astype(np.int32)])
        self.ids = np.concatenate([self.ids[stay], np.ascontiguousarray(rb[:, 8]).view(np.int64)])
        self.isnew = np.concatenate([self.isnew[stay], np.zeros(len(rb), dtype=np.int32)])
        return nl

    def download(self):
        return dict(PartState=self.PS, PartSpecies=self.spec, GlobalElemID=self.elem, ids=self.ids)


class GpuRank:
    def __init__(self, mesh, prm, rank, world, local):
        self.R = multi.ParticleStepRank(mesh, prm, rank, world, local)
        self.off = self.R.offsets
        self.rank = rank

    def upload(self, PS, spec, elem, ids, isnew=1):
        self.R.step.UploadParticles(PS, spec, elem, IsNewPart=np.full(len(spec), isnew, dtype=np.int32), ids=ids)

    def set_field(self, E):
        self.R.step.SetField(np.ascontiguousarray(E[int(self.off[self.rank]):int(self.off[self.rank + 1])]))

    def deposition(self):
        return self.R.Deposition()

    def push_track(self, dt):
        return self.R.PushAndTrack(dt)

    def download(self):
        return self.R.step.DownloadParticles()


def reference_case(a, rank, world, local):
    """The reference's own multi-rank criterion on its own data: NIG_tracking_DSMC/{periodic,ANSA_box} run with MPI = 1,2,5,10 /
    1,2 (command_line.ini) and must reproduce the committed PartInt for every rank count.  Push + TriaTracking + migration over
    the element partition, checked against the state file the reference wrote (tests/test_reference_tracking.py)."""
    import test_reference_tracking as trt
    build = trt.periodic_case if a.case == "periodic_ref" else trt.ansa_case
    mesh, prm, PD0, elem0, PD1, elem1, dt, nsteps = build(hm.TRIATRACKING)
    PS, spec = np.ascontiguousarray(PD0[:, :6]), PD0[:, 6].astype(np.int32)
    ids = np.arange(len(spec), dtype=np.int64)
    off = hm.partition(mesh, world)
    mine = (elem0 > off[rank]) & (elem0 <= off[rank + 1])
    eng = GpuRank(mesh, prm, rank, world, local) if a.engine == "gpu" else OracleRank(mesh, prm, rank, world)
    eng.upload(PS[mine], spec[mine], elem0[mine], ids[mine], isnew=0)
    eng.set_field(np.zeros((mesh.nElems, 2, 2, 2, 3)))
    migrated = 0
    for _ in range(nsteps):
        before = set(eng.download()["ids"].tolist()) if a.engine == "oracle" else None
        assert eng.push_track(dt) == 0
        if before is not None:
            migrated += len(set(eng.download()["ids"].tolist()) - before)
    d = eng.download()
    own = (d["GlobalElemID"] > off[rank]) & (d["GlobalElemID"] <= off[rank + 1])
    assert own.all(), "rank holds particles of elements it does not own"
    parts = [None] * world if rank == 0 else None
    dist.gather_object((d["ids"], d["PartState"], d["GlobalElemID"], migrated), parts, dst=0)
    if rank == 0:
        allid = np.concatenate([p[0] for p in parts])
        o = np.argsort(allid)
        assert np.array_equal(allid[o], ids), "particles lost or duplicated in migration"
        PSe = np.concatenate([p[1] for p in parts])[o]
        ele = np.concatenate([p[2] for p in parts])[o]
        if a.case == "periodic_ref":
            trt.check_periodic(PSe, ele, PD1, elem1, mesh.nElems)
        else:
            trt.check_ansa(PSe, ele, PD0, PD1, elem1, mesh.nElems)
        nmig = sum(p[3] for p in parts)
        assert a.engine != "oracle" or nmig > len(spec), "the case does not exercise the migration"
        print("MULTI_OK engine=%s world=%d case=%s migrations=%d" % (a.engine, world, a.case, nmig))
    if a.engine == "gpu":
        eng.R.close()
    dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--engine", default="oracle")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--particles", type=int, default=12000)
    ap.add_argument("--depo", default="cvwm", choices=["cvwm", "sf", "cc"])
    ap.add_argument("--case", default="synthetic", choices=["synthetic", "periodic_ref", "ansa_ref"])
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.engine == "gpu":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    if a.case != "synthetic":
        return reference_case(a, rank, world, local)

    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (4, 3, 6), 2)
    def params():
        if a.depo == "cvwm":
            return cases.electron_params()
        q = cases.electron_params(DepositionType=DEPO_SF if a.depo == "sf" else DEPO_SF_CC)
        hm.shape_function_setup(mesh, q, 0.3, 2, dim_sf=3)
        return q
    if a.depo != "cvwm":
        hm.add_fibgm(mesh)
    prm = params()
    dt = 2e-8   # keeps the Maxwellian tail far below c (gamma stays finite)
    PS, spec = cases.uniform_plasma(mesh, a.particles, seed=77, vth_cells=0.45, dt=dt)   # identical on every rank
    elem = hm.cartesian_locate(mesh, PS[:, :3])
    ids = np.arange(len(spec), dtype=np.int64)
    E = cases.smooth_field(mesh, amp=1e-4)
    off = hm.partition(mesh, world)
    mine = (elem > off[rank]) & (elem <= off[rank + 1])
    eng = GpuRank(mesh, prm, rank, world, local) if a.engine == "gpu" else OracleRank(mesh, prm, rank, world)
    eng.upload(PS[mine], spec[mine], elem[mine], ids[mine])
    eng.set_field(E)

    # single-rank reference on rank 0
    if rank == 0:
        ref = Oracle(mesh, params())
        PSr, elr = PS.copy(), elem.copy()
        inside = np.ones(len(spec), dtype=np.int32)
        isnew = np.ones(len(spec), dtype=np.int32)
    ok = True
    for it in range(a.steps):
        PSrc, NS = eng.deposition()
        lost = eng.push_track(dt)
        d = eng.download()
        # gather everything on rank 0 (variable sizes -> object gather)
        parts = [None] * world if rank == 0 else None
        dist.gather_object((d["ids"], d["PartState"], d["GlobalElemID"], NS, PSrc), parts, dst=0)
        if rank == 0:
            PSo, NSo = ref.deposit(PSr, spec, elr, inside)
            ref.push_track(dt, PSr, spec, elr, inside, isnew, E)
            allid = np.concatenate([p[0] for p in parts])
            allps = np.concatenate([p[1] for p in parts])
            allel = np.concatenate([p[2] for p in parts])
            o = np.argsort(allid)
            assert np.array_equal(allid[o], np.arange(len(spec))), "particles lost or duplicated in migration"
            assert np.array_equal(allel[o], elr), "element ownership differs from the single-rank run"
            ex = np.abs(allps[o] - PSr).max() / np.abs(PSr).max()
            assert ex <= 1e-12, ex
            for r, p in enumerate(parts):   # after the halo sum every rank holds the NodeSource of the nodes of its own elements
                mine_nodes = np.unique(mesh.NodeInfo[mesh.ElemNodeID[int(off[r]):int(off[r + 1])].reshape(-1) - 1] - 1)
                for c in range(4 if p[3] is not None else 0):
                    en = np.abs(p[3][mine_nodes, c] - NSo[mine_nodes, c]).max() / max(np.abs(NSo[:, c]).max(), 1e-300)
                    assert en <= 1e-12, (r, c, en)
                sl = slice(int(off[r]), int(off[r + 1]))
                es = np.abs(p[4] - PSo[sl]).max() / np.abs(PSo).max()
                assert es <= 1e-12, es
                own = (p[2] > off[r]) & (p[2] <= off[r + 1])
                assert own.all(), "rank holds particles of elements it does not own"
    if rank == 0:
        print("MULTI_OK engine=%s world=%d steps=%d depo=%s" % (a.engine, world, a.steps, a.depo))
    if a.engine == "gpu":
        eng.R.close()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
