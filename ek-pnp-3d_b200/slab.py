"""Multi-GPU path: x-slab decomposition of the coupled step.

One slab = one `Simulation(slab=(rank, nranks))` handle = one process/GPU.
This module is the host-side plumbing around the device pieces of
csrc/ek_slab.cu:

  * population halos: after every LBM step the 9 face-crossing populations of
    each set travel to the two x-neighbours (phase A after an even A-A step,
    phase B after an odd one) -- NCCL send/recv between ring neighbours;
  * phi halo: one column per face for the fused E = -grad(phi);
  * Poisson: local real FFT along y, all-to-all transpose to full-x pencils,
    complex FFT along x, the hand-written z-solve (ek_zsolve_columns), and back.
    The transforms are cuFFT (through torch.fft); the transposes are NCCL
    all-to-alls.

`Comm` abstracts the transport: `DistComm` is torch.distributed (NCCL, one
slab per process); `LocalComm` keeps all slabs of a group in ONE process on one
GPU and replaces the collectives by copies, so that the whole multi-slab
algorithm can be checked against the single-domain run on a single GPU
(tests/test_slab_gpu.py) -- ranks are emulated as data, never as concurrently
waiting kernels.
"""
from __future__ import annotations

import ctypes as C
import json
import time

import numpy as np
import torch


class _DevArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor."""

    def __init__(self, ptr: int, shape, dtype="<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": dtype, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_view(ptr: int, shape, device) -> torch.Tensor:
    return torch.as_tensor(_DevArray(ptr, shape), device=device)


def partition(NX: int, nranks: int):
    """[(x0, x1)] owned by each rank; NX must divide evenly."""
    if NX % nranks:
        raise ValueError(f"NX={NX} is not divisible by {nranks} ranks")
    w = NX // nranks
    return [(r * w, (r + 1) * w) for r in range(nranks)]


def ky_chunks(NY: int, nranks: int):
    """Rows of the y half spectrum (NY//2+1) owned by each rank in the
    transposed layout: equal chunks of ceil(NYH/nranks), the tail is padding."""
    nyh = NY // 2 + 1
    kyl = -(-nyh // nranks)
    return nyh, kyl


# ---------------------------------------------------------------------------
# tensor algebra of the distributed Poisson stage (device independent)
# ---------------------------------------------------------------------------
def y_forward(g: torch.Tensor, P: int, kyl: int) -> torch.Tensor:
    """g (M, NY, NXl) real: FFT along y, rows split into P chunks of kyl -> (P, M, kyl, NXl) complex"""
    M, NY, NXl = g.shape
    G = torch.fft.rfft(g, dim=1)
    send = torch.zeros((M, P * kyl, NXl), dtype=torch.complex128, device=g.device)
    send[:, :NY // 2 + 1, :] = G
    return send.view(M, P, kyl, NXl).permute(1, 0, 2, 3).contiguous()


def x_forward(recv: torch.Tensor) -> torch.Tensor:
    """recv (P, M, kyl, NXl): every rank's x block of my ky chunk -> full-x spectrum (M, kyl, NX)"""
    P, M, kyl, NXl = recv.shape
    X = recv.permute(1, 2, 0, 3).reshape(M, kyl, P * NXl)
    return torch.fft.fft(X, dim=2).contiguous()


def x_backward(X: torch.Tensor, P: int) -> torch.Tensor:
    """inverse (unnormalised) x FFT, split back into x blocks -> (P, M, kyl, NXl)"""
    M, kyl, NX = X.shape
    X = torch.fft.ifft(X, dim=2, norm="forward")
    return X.view(M, kyl, P, NX // P).permute(2, 0, 1, 3).contiguous()


def y_backward(recv: torch.Tensor, NY: int) -> torch.Tensor:
    """recv (P, M, kyl, NXl): every ky chunk of my x block -> (M, NY, NXl) real (unnormalised)"""
    P, M, kyl, NXl = recv.shape
    G = recv.permute(1, 0, 2, 3).reshape(M, P * kyl, NXl)[:, :NY // 2 + 1, :]
    return torch.fft.irfft(G, n=NY, dim=1, norm="forward")


# ---------------------------------------------------------------------------
# transports
# ---------------------------------------------------------------------------
class LocalComm:
    """All slabs live in this process: collectives become copies."""

    def __init__(self, nranks: int):
        self.nranks = nranks
        self.local_ranks = list(range(nranks))
        _blocking(self)

    def neighbor_exchange_start(self, to_left, to_right, from_left, from_right):
        P = self.nranks
        for r in range(P):
            from_left[r].copy_(to_right[(r - 1) % P])
            from_right[r].copy_(to_left[(r + 1) % P])
        return None

    def neighbor_exchange_finish(self, handle):
        pass

    def all_to_all_start(self, send):
        P = self.nranks
        return [torch.stack([send[p][r] for p in range(P)]) for r in range(P)]

    def all_to_all_finish(self, handle):
        return handle

    def barrier(self):
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        return x


class DistComm:
    """One slab per process, torch.distributed (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, dist):
        self.dist = dist
        self.rank = dist.get_rank()
        self.nranks = dist.get_world_size()
        self.local_ranks = [self.rank]
        _blocking(self)

    def neighbor_exchange_start(self, to_left, to_right, from_left, from_right):
        """post the ring send/recv; the transfer runs on NCCL's stream while the
        caller keeps launching kernels, until neighbor_exchange_finish()"""
        d, P, r = self.dist, self.nranks, self.rank
        left, right = (r - 1) % P, (r + 1) % P
        if P == 1:
            from_left[0].copy_(to_right[0])
            from_right[0].copy_(to_left[0])
            return []
        # order matters when left == right (P = 2): the peer's first send (its
        # to_right) is my from_left, its second (to_left) my from_right
        ops = [d.P2POp(d.isend, to_right[0], right), d.P2POp(d.isend, to_left[0], left),
               d.P2POp(d.irecv, from_left[0], left), d.P2POp(d.irecv, from_right[0], right)]
        return d.batch_isend_irecv(ops)

    def neighbor_exchange_finish(self, handle):
        for req in handle:
            req.wait()

    def all_to_all_start(self, send):
        recv = torch.empty_like(send[0])
        if self.nranks == 1:
            recv.copy_(send[0])
            return (None, recv)
        return (self.dist.all_to_all_single(recv, send[0], async_op=True), recv)

    def all_to_all_finish(self, handle):
        work, recv = handle
        if work is not None:
            work.wait()
        return [recv]

    def barrier(self):
        self.dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def _blocking(comm):
    """blocking forms, shared by both transports"""
    def neighbor_exchange(to_left, to_right, from_left, from_right):
        comm.neighbor_exchange_finish(comm.neighbor_exchange_start(to_left, to_right, from_left, from_right))

    def all_to_all(send):
        return comm.all_to_all_finish(comm.all_to_all_start(send))
    comm.neighbor_exchange = neighbor_exchange
    comm.all_to_all = all_to_all
    return comm


# ---------------------------------------------------------------------------
# one slab
# ---------------------------------------------------------------------------
class Slab:
    def __init__(self, ek, params, rank: int, nranks: int, device: int, zchunk=None):
        self.ek = ek
        self.rank, self.nranks = rank, nranks
        self.sim = ek.Simulation(params, device=device, slab=(rank, nranks), zchunk=zchunk)
        self.L, self.h = self.sim.L, self.sim.h
        self.dev = torch.device("cuda", device)
        self.NXg = params.NX
        self.NX, self.NY, self.NZ = self.sim.p.NX, self.sim.p.NY, self.sim.p.NZ
        self.M = self.NZ - 2
        with torch.cuda.device(self.dev):
            self.sim._ck(self.L.ek_set_stream(self.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)), "ek_set_stream")
        self.sim._ck(self.L.ek_ensure_allocated(self.h), "ek_ensure_allocated")
        self.PX = self.L.ek_row_pitch(self.h)
        p = C.c_void_p()
        self.sim._ck(self.L.ek_dq_ptr(self.h, C.byref(p)), "ek_dq_ptr")
        self.dq = device_view(p.value, (self.NZ, self.NY, self.PX), self.dev)
        self.sim._ck(self.L.ek_field_ptr(self.h, ek.FIELDS.index("phi"), C.byref(p)), "ek_field_ptr")
        self.phi = device_view(p.value, (self.NZ, self.NY, self.PX), self.dev)
        nh = self.L.ek_halo_doubles(self.h)
        mk = lambda n: torch.empty(n, dtype=torch.float64, device=self.dev)  # noqa: E731
        self.h_to_l, self.h_to_r, self.h_from_l, self.h_from_r = mk(nh), mk(nh), mk(nh), mk(nh)
        np_ = self.NY * self.NZ
        self.p_to_l, self.p_to_r, self.p_from_l, self.p_from_r = mk(np_), mk(np_), mk(np_), mk(np_)
        self.nyh, self.kyl = ky_chunks(self.NY, nranks)

    def ck(self, st, what):
        self.sim._ck(st, what)

    # -- Poisson pieces --------------------------------------------------------
    def poisson_forward_local(self, z0: int = 0, z1: int | None = None) -> torch.Tensor:
        """real FFT along y of interior planes [z0, z1), split by ky chunk: (P, z1-z0, kyl, NX) complex"""
        z1 = self.M if z1 is None else z1
        return y_forward(self.dq[1 + z0:1 + z1, :, :self.NX], self.nranks, self.kyl)

    def poisson_middle(self, recv: torch.Tensor) -> torch.Tensor:
        """recv (P, M, kyl, NXl): my ky chunk, every rank's x block -> x FFT, z-solve, back"""
        X = x_forward(recv)
        self.ck(self.L.ek_zsolve_columns(self.h, C.c_void_p(X.data_ptr()), self.rank * self.kyl, self.kyl),
                "ek_zsolve_columns")
        return x_backward(X, self.nranks)

    def poisson_backward_local(self, recv: torch.Tensor, z0: int = 0, z1: int | None = None, last: bool = True):
        """recv (P, z1-z0, kyl, NX): every ky chunk of my x block -> inverse real FFT along y -> phi"""
        z1 = self.M if z1 is None else z1
        self.phi[1 + z0:1 + z1, :, :self.NX] = y_backward(recv, self.NY)
        if last:
            self.ck(self.L.ek_poisson_finish(self.h, 0), "ek_poisson_finish")

    def zsolve(self, X: torch.Tensor):
        self.ck(self.L.ek_zsolve_columns(self.h, C.c_void_p(X.data_ptr()), self.rank * self.kyl, self.kyl),
                "ek_zsolve_columns")


# ---------------------------------------------------------------------------
# a group of slabs driven together (all of them with LocalComm, one with DistComm)
# ---------------------------------------------------------------------------
class SlabGroup:
    def __init__(self, ek, params, comm, device: int = 0, zchunk=None):
        self.ek, self.comm, self.params = ek, comm, params
        self.nranks = comm.nranks
        self.slabs = [Slab(ek, params, r, self.nranks, device, zchunk) for r in comm.local_ranks]
        self.t = 0.0
        self.zchunks = 4              # pipeline depth of the Poisson transposes
        self.profile = False          # per-phase CUDA-event timing (development aid)
        self._ev = []

    def _mark(self, name):
        if self.profile:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._ev.append((name, e))

    def phase_times(self) -> dict:
        """ms per phase accumulated since the last call (profile=True)"""
        torch.cuda.synchronize()
        out = {}
        for (n0, e0), (n1, e1) in zip(self._ev[:-1], self._ev[1:]):
            if n1 != "begin":
                out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        self._ev = []
        return out

    def close(self):
        for s in self.slabs:
            s.sim.close()

    # -- exchanges ---------------------------------------------------------------
    def halo_exchange_start(self, phase: int):
        for s in self.slabs:
            s.ck(s.L.ek_halo_pack(s.h, phase, C.c_void_p(s.h_to_l.data_ptr()), C.c_void_p(s.h_to_r.data_ptr())), "ek_halo_pack")
        return self.comm.neighbor_exchange_start([s.h_to_l for s in self.slabs], [s.h_to_r for s in self.slabs],
                                                 [s.h_from_l for s in self.slabs], [s.h_from_r for s in self.slabs])

    def halo_exchange_finish(self, phase: int, handle):
        self.comm.neighbor_exchange_finish(handle)
        for s in self.slabs:
            s.ck(s.L.ek_halo_unpack(s.h, phase, C.c_void_p(s.h_from_l.data_ptr()), C.c_void_p(s.h_from_r.data_ptr())), "ek_halo_unpack")

    def halo_exchange(self, phase: int):
        self.halo_exchange_finish(phase, self.halo_exchange_start(phase))

    def phi_halo_exchange(self):
        for s in self.slabs:
            s.ck(s.L.ek_phi_halo_pack(s.h, C.c_void_p(s.p_to_l.data_ptr()), C.c_void_p(s.p_to_r.data_ptr())), "ek_phi_halo_pack")
        self.comm.neighbor_exchange([s.p_to_l for s in self.slabs], [s.p_to_r for s in self.slabs],
                                    [s.p_from_l for s in self.slabs], [s.p_from_r for s in self.slabs])
        for s in self.slabs:
            s.ck(s.L.ek_phi_halo_unpack(s.h, C.c_void_p(s.p_from_l.data_ptr()), C.c_void_p(s.p_from_r.data_ptr())), "ek_phi_halo_unpack")

    def poisson(self):
        """the distributed fast_Poisson(): dq -> phi (interior, walls, ghost columns).
        The planes are processed in `self.zchunks` chunks so that the all-to-all of
        one chunk travels while the next chunk is being transformed."""
        M = self.slabs[0].M
        K = max(1, min(self.zchunks, M))
        bounds = [(M * k // K, M * (k + 1) // K) for k in range(K)]
        P = self.nranks
        # y FFT + first transpose, chunk by chunk
        pending = []
        for (a, b) in bounds:
            send = [s.poisson_forward_local(a, b) for s in self.slabs]
            pending.append(self.comm.all_to_all_start(send))
        self._mark("poisson_y_fft_pack")
        Xs = [torch.empty((M, s.kyl, s.NXg), dtype=torch.complex128, device=s.dev) for s in self.slabs]
        for (a, b), hnd in zip(bounds, pending):
            recv = self.comm.all_to_all_finish(hnd)
            for s, X, r in zip(self.slabs, Xs, recv):
                X[a:b] = x_forward(r)
        self._mark("poisson_all_to_all_1_x_fft")
        for s, X in zip(self.slabs, Xs):
            s.zsolve(X)
        self._mark("poisson_zsolve")
        pending = []
        for (a, b) in bounds:
            send = [x_backward(X[a:b], P) for X in Xs]
            pending.append(self.comm.all_to_all_start(send))
        self._mark("poisson_x_ifft_pack")
        for k, ((a, b), hnd) in enumerate(zip(bounds, pending)):
            recv = self.comm.all_to_all_finish(hnd)
            for s, r in zip(self.slabs, recv):
                s.poisson_backward_local(r, a, b, last=(k == K - 1))
        self._mark("poisson_all_to_all_2_y_ifft")
        self.phi_halo_exchange()
        self._mark("phi_halo")

    # -- the reference's call sequence ---------------------------------------------
    def initialization(self):
        """initialization() of the reference (LBM.cu:68-146) on the decomposed domain"""
        iters = self.params.pb_iters
        for s in self.slabs:
            s.ck(s.L.ek_init_uniform(s.h), "ek_init_uniform")
        for it in range(iters):
            for s in self.slabs:
                s.ck(s.L.ek_pbe(s.h), "ek_pbe")
            self.poisson()
            if it == iters - 1:
                for s in self.slabs:
                    s.ck(s.L.ek_compute_efield(s.h), "ek_compute_efield")   # E of the un-relaxed phi
            for s in self.slabs:
                s.ck(s.L.ek_pbe_relax(s.h), "ek_pbe_relax")
        for s in self.slabs:
            s.ck(s.L.ek_mark_fields_ready(s.h), "ek_mark_fields_ready")
        self.t = 0.0

    def set_fields(self, fields_global: dict):
        """upload global arrays (NZ, NY, NXglobal): every slab takes its columns"""
        parts = partition(self.params.NX, self.nranks)
        for s in self.slabs:
            x0, x1 = parts[s.rank]
            s.sim.set_fields({k: np.ascontiguousarray(np.asarray(v)[:, :, x0:x1]) for k, v in fields_global.items()})

    def init_equilibrium(self):
        for s in self.slabs:
            s.sim.init_equilibrium()

    def init(self):
        self.initialization()
        self.init_equilibrium()

    def step(self, nsteps: int = 1):
        for i in range(nsteps):
            full = i == nsteps - 1
            parity = self.slabs[0].L.ek_lbm_parity(self.slabs[0].h)
            self._mark("begin")
            for s in self.slabs:
                s.sim.stream_collide_save(full)
            self._mark("lbm")
            # the populations travel while the Poisson stage computes (independent data)
            phase = 0 if parity == 0 else 1
            hnd = self.halo_exchange_start(phase)
            self._mark("population_halo_pack")
            self.poisson()
            self.halo_exchange_finish(phase, hnd)
            self._mark("population_halo_unpack")
            if full:
                for s in self.slabs:
                    s.ck(s.L.ek_compute_efield(s.h), "ek_compute_efield")
        self.t += nsteps * self.params.dt

    # -- data ------------------------------------------------------------------------
    def local_fields(self) -> dict:
        """{rank: {name: array (NZ, NY, NXlocal)}} of the slabs of this process"""
        return {s.rank: s.sim.fields() for s in self.slabs}

    def gather_fields(self) -> dict:
        """global arrays (LocalComm only: every slab is here)"""
        parts = self.local_fields()
        return {k: np.concatenate([parts[r][k] for r in sorted(parts)], axis=2) for k in self.ek.FIELDS}

    def gather_populations(self, s_id: int) -> np.ndarray:
        return np.concatenate([s.sim.populations(s_id) for s in sorted(self.slabs, key=lambda q: q.rank)], axis=3)


# ---------------------------------------------------------------------------
# bench.py leg for N > 1 (torchrun, one rank per GPU)
# ---------------------------------------------------------------------------
def bench_slabs(ek, dist, args, w, wl, local_rank):
    from bench import B_ALG_STEP, ClockSampler, measured_peak  # noqa: PLC0415
    comm = DistComm(dist)
    NX, NY, NZ = w["NX"], w["NY"], w["NZ"]
    p = ek.default_params(NX=NX, NY=NY, NZ=NZ, pb_iters=args.pb_iters, **w["over"])
    grp = SlabGroup(ek, p, comm, device=local_rank, zchunk=args.zchunk)
    t0 = time.time()
    grp.init()
    comm.barrier()
    init_s = time.time() - t0
    sampler = ClockSampler(local_rank)
    sampler.start()
    grp.step(args.warmup)
    comm.barrier()
    time.sleep(0.5)
    grp.step(args.warmup)
    comm.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record()
    grp.step(args.steps)
    e1.record()
    torch.cuda.synchronize()
    ms_local = e0.elapsed_time(e1)
    sampler.window(tw0, time.time())
    clocks = sampler.stop()
    comm.barrier()
    ms = comm.max_over_ranks(ms_local)
    cells = NX * NY * NZ
    mlups = cells * args.steps / (ms * 1e-3) / 1e6
    peak, peak_src = measured_peak()
    launches = sum(int(s.sim.counter("kernel_launches")) for s in grp.slabs)
    grp.profile = True
    grp.step(max(4, args.warmup))
    phases = {k: round(v / max(4, args.warmup), 4) for k, v in grp.phase_times().items()}
    grp.close()
    step_gbs = mlups * 1e6 * B_ALG_STEP / 1e9
    return {"metric": "coupled_step_mlups", "value": round(mlups, 2), "unit": "MLUPS", "n_gpus": comm.nranks,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak" if w.get("weak") else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "grid": [NX, NY, NZ], "stream_mode": "aa", "zchunk": args.zchunk,
                       "parallelism": f"x-slabs x{comm.nranks}: NCCL halo send/recv + all-to-all Poisson transposes",
                       "cells_per_gpu": cells // comm.nranks, "init": "reference start-up (PB iterations) %.2f s" % init_s,
                       "l2": "per-GPU working set >> 126 MB L2", "phase_ms_rank0": phases},
            "roofline": {"bound": "hbm", "achieved": round(step_gbs / comm.nranks, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(step_gbs / comm.nranks / peak, 4), "peak_source": peak_src,
                         "note": "whole coupled step per GPU at 1760 B/cell (kernel split is reported at N=1)",
                         "traffic": None},
            "cpu_baseline": None,
            "e2e": None, "gpu_launches": launches, "clocks": clocks,
            "hbm_gbs_step": round(step_gbs, 1)}
