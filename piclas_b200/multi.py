"""Multi-GPU driver of the particle step: one process per GPU, element partition as PICLas' MPI decomposition.

What is exchanged, all through torch.distributed (NCCL on GPUs, gloo in the CPU tests):
  * particle migration after tracking (reference particle_mpi.f90:202-1024): per-destination counts, then the
    AoS particle messages (PartCommSize doubles each) with one all-to-all-v;
  * cell_volweight_mean node halo (pic_depo_method.f90:565-673): the rank-local node sums are added over ranks
    before the division by NodeVolume.
The element partition is the reference's equal split of the element range (loadbalance/loaddistribution.f90:362-369),
written into ElemInfo(ELEM_RANK,:) by hostmesh.partition().

The device work is behind the C ABI (piclas_gpu_exchange_* / piclas_gpu_nodesource_device / piclas_gpu_deposit_finish);
this module is transport plumbing only and is shared by the CPU (gloo) tests, which drive it with a CPU engine.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from . import hostmesh as hm
from .abi import DEPO_CVWM, Params


class _DevPtr:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy, no ownership)."""

    def __init__(self, ptr, nelem, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(nelem),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_tensor(ptr, nelem, device, typestr="<f8"):
    if nelem == 0 or not ptr:
        return torch.empty(0, dtype=torch.int64 if typestr == "<i8" else torch.float64, device=device)
    return torch.as_tensor(_DevPtr(ptr, nelem, typestr), device=device)


# ---- transport (works for NCCL/cuda tensors and gloo/cpu tensors) -------------------------------------------------------
def exchange_counts(send_counts, device, group=None):
    """IRecvNbOfParticles / SendNbOfParticles (particle_mpi.f90:202-342): counts to / from every rank."""
    world = dist.get_world_size(group)
    s = torch.tensor(list(send_counts), dtype=torch.int64, device=device)
    r = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(r, s, group=group)
    return [int(v) for v in r.tolist()]


def exchange_particles(send_buf, send_counts, recv_buf, recv_counts, comm_size, group=None):
    """MPIParticleSend / MPIParticleRecv (particle_mpi.f90:344-1024): flat double buffers, comm_size doubles/particle."""
    ins = [int(c) * comm_size for c in send_counts]
    outs = [int(c) * comm_size for c in recv_counts]
    dist.all_to_all_single(recv_buf, send_buf, output_split_sizes=outs, input_split_sizes=ins, group=group)


def halo_sum(node_tensor, group=None):
    """cell_volweight_mean node halo (pic_depo_method.f90:565-673) as a sum over all ranks."""
    dist.all_reduce(node_tensor, op=dist.ReduceOp.SUM, group=group)


def exchange_sf_halo(send_buf, send_elems, recv_buf, recv_elems, doubles_per_elem, group=None):
    """Shape-function DOF halo (pic_depo_method.f90:940-996, ShapeMapping Send/RecvBuffer): every rank sends the
    PartSource blocks its particles formed on elements of other ranks; blocks are ordered by owner rank."""
    ins = [int(c) * doubles_per_elem for c in send_elems]
    outs = [int(c) * doubles_per_elem for c in recv_elems]
    dist.all_to_all_single(recv_buf, send_buf, output_split_sizes=outs, input_split_sizes=ins, group=group)


# ---- the GPU rank --------------------------------------------------------------------------------------------------------
class ParticleStepRank:
    """ParticleStep of one rank + the exchanges that follow its operators (GPU, NCCL)."""

    def __init__(self, mesh, params: Params, rank, world, local_rank=0, group=None):
        from .particle_step import ParticleStep
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device("cuda", local_rank)
        off = hm.partition(mesh, world)
        self.offsets = off
        params.myRank, params.nRanks, params.device = rank, world, local_rank
        self.step = ParticleStep(mesh, params, offsetElem=int(off[rank]), nElems=int(off[rank + 1] - off[rank]))
        self.lib = self.step.lib
        self.mesh = mesh
        self.migrated = 0
        # the library works on torch's current stream: its kernels and the NCCL collectives issued below are then ordered on the
        # device, no host synchronisation between them
        self._check(self.lib.piclas_gpu_set_stream(C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        # PICLAS_MULTI_TIMING=1: CUDA-event marks between the calls of a step (bench diagnostics; events only, no synchronisation)
        self._marks = [] if os.environ.get("PICLAS_MULTI_TIMING") else None
        self._wrapped = {}                                                   # (device pointer, capacity) -> tensor over the whole buffer
        self._rcnt = torch.zeros(world, dtype=torch.int64, device=self.device)

    def _mark(self, name):
        if self._marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._marks.append((name, ev, time.perf_counter()))

    def timing_report(self):
        """Average milliseconds between consecutive marks, on the device (events) and on the host (perf_counter)."""
        if not self._marks:
            return {}
        torch.cuda.synchronize()
        acc, cnt = {}, {}
        for (n0, e0, h0), (n1, e1, h1) in zip(self._marks[:-1], self._marks[1:]):
            k = n0 + " -> " + n1
            a = acc.setdefault(k, [0.0, 0.0])
            a[0] += e0.elapsed_time(e1)
            a[1] += 1e3 * (h1 - h0)
            cnt[k] = cnt.get(k, 0) + 1
        return {k: (round(v[0] / cnt[k], 3), round(v[1] / cnt[k], 3)) for k, v in acc.items()}

    def _buffer(self, ptr, cap, typestr="<f8"):
        """Tensor over a whole device buffer of the library, wrapped once (wrapping costs tens of microseconds) and sliced per step."""
        key = (int(ptr or 0), int(cap), typestr)
        t = self._wrapped.get(key)
        if t is None:
            if len(self._wrapped) > 64:
                self._wrapped.clear()
            t = device_tensor(ptr, cap, self.device, typestr)
            self._wrapped[key] = t
        return t

    def close(self):
        self.step.close()

    def _check(self, rc):
        self.step._check(rc)

    def exchange(self):
        world = self.world
        cs = C.c_int32(0)
        nsend = (C.c_int64 * world)()
        sp = C.c_void_p(0)
        self._mark("exchange_info")
        self._check(self.lib.piclas_gpu_exchange_info(C.byref(cs), nsend, C.byref(sp)))
        self._mark("counts")
        send_counts = [int(nsend[r]) for r in range(world)]
        # counts to / from every rank (IRecvNbOfParticles / SendNbOfParticles): the library left them on the device as well
        cp = C.c_void_p(0)
        scap, rcap = C.c_int64(0), C.c_int64(0)
        self._check(self.lib.piclas_gpu_exchange_device_info(C.byref(cp), C.byref(scap), C.byref(rcap)))
        dist.all_to_all_single(self._rcnt, self._buffer(cp.value, world, "<i8"), group=self.group)
        recv_counts = [int(v) for v in self._rcnt.tolist()]
        self._mark("messages")
        nrecv = sum(recv_counts)
        rp = C.c_void_p(0)
        self._check(self.lib.piclas_gpu_exchange_recv_buffer(C.c_int64(nrecv), C.byref(rp)))
        self._check(self.lib.piclas_gpu_exchange_device_info(C.byref(cp), C.byref(scap), C.byref(rcap)))
        nsd, nrd = sum(send_counts) * cs.value, nrecv * cs.value
        sbuf = self._buffer(sp.value, scap.value)[:nsd] if nsd else torch.empty(0, dtype=torch.float64, device=self.device)
        rbuf = self._buffer(rp.value, rcap.value)[:nrd] if nrd else torch.empty(0, dtype=torch.float64, device=self.device)
        exchange_particles(sbuf, send_counts, rbuf, recv_counts, cs.value, self.group)
        self._mark("exchange_finish")
        self._check(self.lib.piclas_gpu_exchange_finish(C.c_int64(nrecv)))
        self._mark("step_end")
        self.migrated = sum(send_counts)
        return sum(send_counts), nrecv

    def PushAndTrack(self, dt, it=0):
        self._mark("push_track")
        lost = self.step.PushAndTrack(dt, it)
        self.exchange()
        return lost

    def Deposition(self, want_partsource=True, want_nodesource=True, out_partsource=None):
        from .particle_step import _f
        if self.step.params.DepositionType != DEPO_CVWM:                            # shape functions: DOF halo
            self._check(self.lib.piclas_gpu_deposit(_f(None), _f(None)))
            ns = (C.c_int64 * self.world)()
            nr = (C.c_int64 * self.world)()
            dpe = C.c_int32(0)
            sp, rp = C.c_void_p(0), C.c_void_p(0)
            self._check(self.lib.piclas_gpu_sf_halo_info(ns, nr, C.byref(dpe), C.byref(sp), C.byref(rp)))
            ns, nr = [int(v) for v in ns], [int(v) for v in nr]
            sbuf = device_tensor(sp.value, sum(ns) * dpe.value, self.device)
            rbuf = device_tensor(rp.value, sum(nr) * dpe.value, self.device)
            exchange_sf_halo(sbuf, ns, rbuf, nr, dpe.value, self.group)
            PS = out_partsource if out_partsource is not None else (np.empty(self.step._ps_shape) if want_partsource else None)
            self._check(self.lib.piclas_gpu_deposit_finish(_f(PS), _f(None)))
            return PS, None
        self._mark("deposit")
        self._check(self.lib.piclas_gpu_deposit(_f(None), _f(None)))       # rank-local node sums
        self._mark("node_halo")
        # node halo on the compact list of shared nodes: all-gather, then the library adds the ranks' parts in rank order
        nd = C.c_int64(0)
        sp, rp = C.c_void_p(0), C.c_void_p(0)
        self._check(self.lib.piclas_gpu_node_halo_info(C.byref(nd), C.byref(sp), C.byref(rp)))
        if nd.value > 0:
            dist.all_gather_into_tensor(self._buffer(rp.value, nd.value * self.world), self._buffer(sp.value, nd.value), group=self.group)
        PS = out_partsource if out_partsource is not None else (np.empty(self.step._ps_shape) if want_partsource else None)
        NS = np.empty((self.mesh.nUniqueNodes, 4)) if want_nodesource else None
        self._mark("deposit_finish")
        self._check(self.lib.piclas_gpu_deposit_finish(_f(PS), _f(NS)))
        return PS, NS


# ---- bench.py --gpus N (N > 1) -------------------------------------------------------------------------------------------
def run_bench_multi(args, rank, world, local):
    import bench as B
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    parity = B.small_parity_check(rank, world, local, args.arithmetic) if not getattr(args, "no_checks", False) else None
    mesh, E, dt, vth = B.workload(args.nelem, args.N, kick=args.kick)
    n_total = int(args.particles)
    prm = Params(ChargeIC=(-B.QE,), MassIC=(B.ME,), MacroParticleFactor=(1.0e3,), device=local, arithmetic=args.arithmetic)
    off = hm.partition(mesh, world)
    n_loc_elems = int(off[rank + 1] - off[rank])
    # particles of this rank: uniform in its element range (z-slabs for the i-fastest element order)
    n_loc = n_total // world + (1 if rank < n_total % world else 0)
    prm.maxParticleNumber = int(n_loc * 1.15) + 4096
    R = ParticleStepRank(mesh, prm, rank, world, local)
    rng = np.random.default_rng(args.seed + 1000 * rank)
    ne = args.nelem
    layer = ne * ne
    done = 0
    chunk = 10_000_000
    while done < n_loc:
        m = min(chunk, n_loc - done)
        # pick elements uniformly from the local range, then a uniform point inside the element
        el = rng.integers(int(off[rank]), int(off[rank + 1]), m)
        ijk = np.stack([el % ne, (el // ne) % ne, el // layer], axis=1)
        x = (ijk + rng.random((m, 3))) / ne
        v = rng.normal(0.0, vth, (m, 3))
        PS = np.ascontiguousarray(np.concatenate([x, v], axis=1))
        ijk2 = np.minimum((x * ne).astype(np.int64), ne - 1)
        elem = (1 + ijk2[:, 0] + ne * (ijk2[:, 1] + ne * ijk2[:, 2])).astype(np.int32)
        elem = np.clip(elem, int(off[rank]) + 1, int(off[rank + 1])).astype(np.int32)
        R.step.UploadParticles(PS, np.ones(m, dtype=np.int32), elem, append=done > 0)
        done += m
    E_loc = np.ascontiguousarray(E[int(off[rank]):int(off[rank + 1])])
    R.step.SetField(E_loc)

    def one_step():
        R.Deposition(want_partsource=False, want_nodesource=False)
        ph_d = R.step.PhaseTiming().copy()
        _, nl_d = R.step.LastTiming()
        R.PushAndTrack(dt)
        ph_p = R.step.PhaseTiming().copy()
        _, nl_p = R.step.LastTiming()
        return nl_d + nl_p, np.array([ph_d[0], ph_d[1], ph_p[2], ph_p[3]])

    for _ in range(args.warmup):
        one_step()
    sampler = B.ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    dist.barrier()
    torch.cuda.synchronize()
    if R._marks is not None:
        R._marks.clear()
    t0 = time.perf_counter()
    launches, phases, migrated = 0, np.zeros(4), 0
    for _ in range(args.steps):
        a, b = one_step()
        launches += a
        phases += b
        migrated += R.migrated
    torch.cuda.synchronize()
    dist.barrier()
    wall = time.perf_counter() - t0
    if R._marks is not None:
        rep = R.timing_report()
        R._marks = None
        if rank == 0 or rank == world - 1:
            print("[multi timing rank %d] (device ms, host ms) " % rank + json.dumps(rep), file=sys.stderr)
    clocks = sampler.stop() if sampler else None
    tmax = torch.tensor([wall], dtype=torch.float64, device="cuda")
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([R.step.NumParticles(), migrated, launches], dtype=torch.int64, device="cuda")
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    phs = torch.tensor(phases / args.steps, dtype=torch.float64, device="cuda")
    dist.all_reduce(phs, op=dist.ReduceOp.MAX)
    wall = float(tmax.item())

    # end to end with host buffers (each rank moves its own E / PartSource slices)
    e2e = None
    if not args.no_e2e:
        n1 = args.N + 1
        E_pin = torch.from_numpy(E_loc).pin_memory()
        PS_pin = torch.empty((n_loc_elems, n1, n1, n1, 4), dtype=torch.float64).pin_memory()
        rho_pin = torch.empty((n_loc_elems, n1, n1, n1), dtype=torch.float64).pin_memory()
        PS_h, rho_h = PS_pin.numpy(), rho_pin.numpy()
        R.Deposition(want_partsource=False, want_nodesource=False)          # untimed: first ChargeDensity call allocates
        R.step.ChargeDensity(out=rho_h)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):                                     # as bench.py's single-GPU loop
            R.step.PartSourceWait()
            R.Deposition(want_partsource=False, want_nodesource=False)
            R.step.ChargeDensity(out=rho_h)
            R.step.PartSourceAsync(PS_h)
            R.step.SetField(E_pin.numpy())
            R.PushAndTrack(dt)
        R.step.PartSourceWait()
        torch.cuda.synchronize()
        dist.barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total * args.e2e_steps / float(te.item()), "unit": "particle-steps/s",
               "h2d_bytes_per_step": int(E.nbytes), "d2h_bytes_per_step": int(mesh.nElems * n1 ** 3 * 5 * 8),
               "steps": args.e2e_steps, "ms_per_step": 1e3 * float(te.item()) / args.e2e_steps,
               "note": "host E in; charge density out before the push (HDG input), whole PartSource out on a copy stream beside the push"}
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):                                     # every copy in line, as round 1 measured it
            R.Deposition(out_partsource=PS_h, want_nodesource=False)
            R.step.SetField(E_pin.numpy())
            R.PushAndTrack(dt)
        torch.cuda.synchronize()
        dist.barrier()
        ts = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e["serial"] = {"value": n_total * args.e2e_steps / float(ts.item()), "ms_per_step": 1e3 * float(ts.item()) / args.e2e_steps,
                         "h2d_bytes_per_step": int(E.nbytes), "d2h_bytes_per_step": int(mesh.nElems * n1 ** 3 * 4 * 8)}
    checks = B.full_size_checks_multi(R, mesh, n_total, -B.QE * 1.0e3, rank, world)
    if parity is not None:
        checks["parity_small_case"] = parity
        checks["ok"] = bool(checks.get("ok") and parity.get("ok"))
    R.close()
    if rank == 0:
        peak, peak_src = B.hbm_peak()
        ph = phs.tolist()
        t_push = ph[2] * 1e-3
        per_gpu = n_total / world
        achieved = B.ALG_BYTES_PER_PARTICLE_STEP * per_gpu / t_push / 1e9 if t_push > 0 else 0.0
        line = {"metric": "particle-steps/s (interp+push+track+depo)", "value": n_total * args.steps / wall,
                "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": B.config_dict(args, n_total), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(cnt[2].item()), "particles_end": int(cnt[0].item()),
                "migrated_per_step": int(cnt[1].item()) // max(args.steps, 1),
                "roofline": {"bound": "hbm", "kernel": "k_bin_push + k_far_hint + k_far_walk (per GPU, slowest rank)",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": B.measured_traffic(per_gpu)[0],
                             "traffic_source": B.measured_traffic(per_gpu)[1],
                             "peak_source": peak_src,
                             "step_frac": (B.ALG_BYTES_PER_PARTICLE_STEP * per_gpu * args.steps / wall / 1e9) / peak,
                             "phase_ms": {"deposit_particles": ph[0], "deposit_nodes_dofs": ph[1],
                                          "interp_push_track": ph[2], "sort_permute": ph[3]}},
                "checks": checks}
        print(json.dumps(line))
    dist.destroy_process_group()
