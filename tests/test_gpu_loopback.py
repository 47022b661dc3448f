"""Single-GPU loopback of the multi-rank device path (collectives C2-C4).

The N>1 code of libpiclas_gpu.so (emigrant extraction + pack in the reference's message layout, unpack of immigrants, node
halo of cell_volweight_mean, DOF halo of the shape functions) is normally driven by one process per GPU over NCCL
(tests/test_multi_rank.py, needs 2 GPUs).  Here ONE GPU plays every rank in turn and the test itself is the transport:
  pass 1: each rank is initialised as rank r of W, steps, and its device send buffers are copied to the host;
  pass 2: each rank repeats the same (deterministic) step, receives what pass 1 collected for it, finishes the exchange.
What is compared with the single-rank CPU oracle: the emigrant set of every (source, destination) pair and its message content
(PartState(1:6), species, element id as REALs, particle_mpi.f90:472-502), every rank's population after the exchange (ids,
ownership exact; x, v <= 1e-12), and the deposited sources after the halo sums (<= 1e-12).
"""
import ctypes as C

import numpy as np
import pytest

import cases
from oracle_lib import Oracle
from piclas_b200 import hostmesh as hm
from piclas_b200.abi import DEPO_CVWM, DEPO_SF

pytestmark = pytest.mark.gpu
RTOL = 1e-12


def _dev(ptr, n):
    import torch
    from piclas_b200.multi import device_tensor
    return device_tensor(ptr, n, torch.device("cuda", 0))


class LoopRank:
    """One rank of W on cuda:0 through the C ABI (piclas_b200.multi.ParticleStepRank without the transport)."""

    def __init__(self, mesh, prm, rank, world):
        from piclas_b200.particle_step import ParticleStep, _f
        self._f = _f
        self.off = hm.partition(mesh, world)
        prm.myRank, prm.nRanks, prm.device = rank, world, 0
        self.rank, self.world, self.mesh = rank, world, mesh
        self.step = ParticleStep(mesh, prm, offsetElem=int(self.off[rank]), nElems=int(self.off[rank + 1] - self.off[rank]))
        self.lib = self.step.lib

    def close(self):
        self.step.close()

    def send_buffers(self):
        """piclas_gpu_exchange_info -> per destination rank the host copy of the messages [n, PartCommSize]."""
        cs = C.c_int32(0)
        ns = (C.c_int64 * self.world)()
        sp = C.c_void_p(0)
        self.step._check(self.lib.piclas_gpu_exchange_info(C.byref(cs), ns, C.byref(sp)))
        counts = [int(v) for v in ns]
        flat = _dev(sp.value, sum(counts) * cs.value).cpu().numpy().reshape(-1, cs.value) if sum(counts) else np.zeros((0, cs.value))
        out, o = [], 0
        for c in counts:
            out.append(flat[o:o + c].copy())
            o += c
        return out, cs.value

    def receive(self, msgs):
        import torch
        n = sum(len(m) for m in msgs)
        rp = C.c_void_p(0)
        self.step._check(self.lib.piclas_gpu_exchange_recv_buffer(C.c_int64(n), C.byref(rp)))
        if n:
            flat = np.ascontiguousarray(np.concatenate(msgs).reshape(-1))
            _dev(rp.value, flat.size).copy_(torch.from_numpy(flat))
            torch.cuda.synchronize()
        self.step._check(self.lib.piclas_gpu_exchange_finish(C.c_int64(n)))


def _setup(depo, arith):
    mesh = hm.box_mesh([0, 0, 0], [1, 1, 1], (4, 3, 6), 2)
    if depo == "sf":
        hm.add_fibgm(mesh)

    def params():
        if depo == "cvwm":
            return cases.electron_params(arithmetic=arith)
        q = cases.electron_params(arithmetic=arith, DepositionType=DEPO_SF)
        hm.shape_function_setup(mesh, q, 0.3, 2, dim_sf=3)
        return q
    return mesh, params


@pytest.mark.parametrize("arith", [0, 1], ids=["reference-order", "restructured"])
@pytest.mark.parametrize("world", [2, 3])
def test_particle_exchange_loopback(world, arith):
    mesh, params = _setup("cvwm", arith)
    dt = 2e-8
    n = 12000
    PS0, spec0 = cases.uniform_plasma(mesh, n, seed=77, vth_cells=0.45, dt=dt)
    elem0 = hm.cartesian_locate(mesh, PS0[:, :3])
    E = cases.smooth_field(mesh, amp=1e-4)
    off = hm.partition(mesh, world)
    rank_of = np.asarray(mesh.ElemInfo[:, 6])
    orc = Oracle(mesh, params())
    PSo, elo = PS0.copy(), elem0.copy()
    inside = np.ones(n, dtype=np.int32)
    isnew_o = np.ones(n, dtype=np.int32)
    # host copies of every rank's population between the steps (what the Fortran host would hold after a download)
    pop = []
    for r in range(world):
        m = (elem0 > off[r]) & (elem0 <= off[r + 1])
        pop.append(dict(PS=PS0[m].copy(), spec=spec0[m].copy(), elem=elem0[m].copy(), ids=np.nonzero(m)[0].astype(np.int64),
                        isnew=np.ones(int(m.sum()), dtype=np.int32)))
    total_migrated = 0
    for it in range(3):
        rank_before = rank_of[elo - 1].copy()
        orc.push_track(dt, PSo, spec0, elo, inside, isnew_o, E)
        rank_after = rank_of[elo - 1]

        def run(r, finish_with=None):
            R = LoopRank(mesh, params(), r, world)
            p = pop[r]
            R.step.UploadParticles(p["PS"], p["spec"], p["elem"], IsNewPart=p["isnew"], ids=p["ids"])
            R.step.SetField(np.ascontiguousarray(E[int(off[r]):int(off[r + 1])]))
            assert R.step.PushAndTrack(dt, it) == 0
            msgs, cs = R.send_buffers()
            d = None
            if finish_with is not None:
                R.receive(finish_with)
                d = R.step.DownloadParticles()
            R.close()
            return msgs, cs, d

        sent = []
        for r in range(world):                      # pass 1: what every rank sends
            msgs, cs, _ = run(r)
            assert cs == 9                          # PartState(6), species, element, id bits (ids are carried in the tests)
            assert len(msgs[r]) == 0
            for dst in range(world):
                want = np.nonzero((rank_before == r) & (rank_after == dst) & (dst != r))[0]
                m = msgs[dst]
                got_ids = np.ascontiguousarray(m[:, 8]).view(np.int64)
                o = np.argsort(got_ids)
                assert np.array_equal(got_ids[o], want), "emigrant set %d -> %d differs from the oracle" % (r, dst)
                assert np.array_equal(m[o, 6], spec0[want].astype(np.float64))          # REAL(PartSpecies)
                assert np.array_equal(m[o, 7], elo[want].astype(np.float64))            # REAL(PEM%GlobalElemID), bit-exact ownership
                if len(want):
                    assert np.abs(m[o, :6] - PSo[want]).max() <= RTOL * np.abs(PSo).max()
                total_migrated += len(want)
            sent.append(msgs)
        for r in range(world):                      # pass 2: same step, then receive and finish
            inbox = [sent[s][r] for s in range(world)]
            _, _, d = run(r, finish_with=inbox)
            want = np.nonzero(rank_after == r)[0]
            o = np.argsort(d["ids"])
            assert np.array_equal(d["ids"][o], want), "population of rank %d after the exchange differs" % r
            assert np.array_equal(d["GlobalElemID"][o], elo[want])
            assert np.abs(d["PartState"][o] - PSo[want]).max() <= RTOL * np.abs(PSo).max()
            loc = d["GlobalElemID"] - int(off[r]) - 1
            assert (np.diff(loc) >= 0).all(), "device order after the exchange is not sorted by element"
            pop[r] = dict(PS=d["PartState"], spec=d["PartSpecies"], elem=d["GlobalElemID"], ids=d["ids"],
                          isnew=np.zeros(len(d["ids"]), dtype=np.int32))
    assert total_migrated > 500, "the case does not exercise the migration"
    orc.close()


@pytest.mark.parametrize("arith", [0, 1], ids=["reference-order", "restructured"])
@pytest.mark.parametrize("depo", ["cvwm", "sf"])
def test_deposition_halo_loopback(depo, arith):
    """C3 (node halo of cell_volweight_mean) and C4 (DOF halo of the shape function) with one GPU playing both ranks."""
    import torch
    world = 2
    mesh, params = _setup(depo, arith)
    n = 12000
    PS, spec = cases.uniform_plasma(mesh, n, seed=78, vth_cells=0.45, dt=2e-8)
    elem = hm.cartesian_locate(mesh, PS[:, :3])
    off = hm.partition(mesh, world)
    orc = Oracle(mesh, params())
    PSr, NSr = orc.deposit(PS, spec, elem, np.ones(n, dtype=np.int32))
    orc.close()

    def start(r):
        R = LoopRank(mesh, params(), r, world)
        m = (elem > off[r]) & (elem <= off[r + 1])
        R.step.UploadParticles(PS[m], spec[m], elem[m], ids=np.nonzero(m)[0].astype(np.int64))
        R.step._check(R.lib.piclas_gpu_deposit(R._f(None), R._f(None)))
        return R

    if depo == "cvwm":
        S = []
        for r in range(world):
            R = start(r)
            p = C.c_void_p(0)
            R.step._check(R.lib.piclas_gpu_nodesource_device(C.byref(p)))
            S.append(_dev(p.value, mesh.nUniqueNodes * 4).cpu().numpy().copy())
            R.close()
        total = S[0].copy()
        for r in range(1, world):
            total = total + S[r]                       # rank order, as the header documents
        for r in range(world):
            R = start(r)
            p = C.c_void_p(0)
            R.step._check(R.lib.piclas_gpu_nodesource_device(C.byref(p)))
            _dev(p.value, total.size).copy_(torch.from_numpy(total))
            torch.cuda.synchronize()
            PSg = np.empty(R.step._ps_shape)
            NSg = np.empty((mesh.nUniqueNodes, 4))
            R.step._check(R.lib.piclas_gpu_deposit_finish(R._f(PSg), R._f(NSg)))
            R.close()
            sl = slice(int(off[r]), int(off[r + 1]))
            for c in range(4):
                assert np.abs(NSg[:, c] - NSr[:, c]).max() <= RTOL * np.abs(NSr[:, c]).max()
                assert np.abs(PSg[..., c] - PSr[sl][..., c]).max() <= RTOL * np.abs(PSr[..., c]).max()
        return
    # shape function: blocks for elements of the other rank
    sent, dpe = [], None
    for r in range(world):
        R = start(r)
        ns, nr = (C.c_int64 * world)(), (C.c_int64 * world)()
        d = C.c_int32(0)
        sp, rp = C.c_void_p(0), C.c_void_p(0)
        R.step._check(R.lib.piclas_gpu_sf_halo_info(ns, nr, C.byref(d), C.byref(sp), C.byref(rp)))
        dpe = d.value
        cnt = [int(v) for v in ns]
        flat = _dev(sp.value, sum(cnt) * dpe).cpu().numpy().copy() if sum(cnt) else np.zeros(0)
        blocks, o = [], 0
        for c in cnt:
            blocks.append(flat[o:o + c * dpe])
            o += c * dpe
        sent.append(blocks)
        R.close()
    assert sum(b.size for blocks in sent for b in blocks) > 0, "no shape-function halo in this case"
    for r in range(world):
        R = start(r)
        ns, nr = (C.c_int64 * world)(), (C.c_int64 * world)()
        d = C.c_int32(0)
        sp, rp = C.c_void_p(0), C.c_void_p(0)
        R.step._check(R.lib.piclas_gpu_sf_halo_info(ns, nr, C.byref(d), C.byref(sp), C.byref(rp)))
        inbox = np.ascontiguousarray(np.concatenate([sent[s][r] for s in range(world)]))
        assert inbox.size == sum(int(v) for v in nr) * dpe, "send and receive element lists of the two ranks do not mirror each other"
        if inbox.size:
            _dev(rp.value, inbox.size).copy_(torch.from_numpy(inbox))
            torch.cuda.synchronize()
        PSg = np.empty(R.step._ps_shape)
        R.step._check(R.lib.piclas_gpu_deposit_finish(R._f(PSg), R._f(None)))
        R.close()
        sl = slice(int(off[r]), int(off[r + 1]))
        for c in range(4):
            assert np.abs(PSg[..., c] - PSr[sl][..., c]).max() <= RTOL * max(np.abs(PSr[..., c]).max(), 1e-300)
