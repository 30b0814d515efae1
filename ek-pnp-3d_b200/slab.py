"""Multi-GPU path: x-slab decomposition of the coupled step.

One slab = one `Simulation(slab=(rank, nranks))` handle = one process/GPU.
This module is the host-side plumbing around the device pieces of
csrc/ek_slab.cu:

  * population halos: after every LBM step the 9 face-crossing populations of
    each set travel to the two x-neighbours (phase A after an even A-A step,
    phase B after an odd one) -- NCCL send/recv between ring neighbours;
  * phi halo: one column per face for the fused E = -grad(phi);
  * Poisson: the native stage of csrc/ek_slab_poisson.cu (cuFFT transforms,
    hand-written re-blocking and z-solve kernels) chunked along z; this module
    only moves the chunk buffers between ranks (NCCL all-to-all) and pipelines
    chunk k's forward half against the LBM launch of chunk k+1.
    The torch.fft functions below (y_forward ... y_backward) restate the same
    algebra device-independently; they are the CPU tests' model of the stage
    (tests/test_slab_cpu.py, gloo), not the product path.

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


def device_view(ptr: int, shape, device, dtype="<f8") -> torch.Tensor:
    return torch.as_tensor(_DevArray(ptr, shape, dtype), device=device)


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


def x_backward_ghost(X: torch.Tensor, P: int) -> torch.Tensor:
    """x_backward with the two ghost columns of phi inside the transpose (what ek_rank.cu does,
    ek_slab_poisson_scatter_xg): rank i's block carries its NXl columns plus column (i+1)*NXl and column
    i*NXl - 1 (periodic in x) -> (P, M, kyl, NXl + 2).  After y_backward the last two columns are the
    right / left ghost column of phi: no separate phi halo exchange."""
    M, kyl, NX = X.shape
    NXl = NX // P
    X = torch.fft.ifft(X, dim=2, norm="forward")
    out = torch.empty((P, M, kyl, NXl + 2), dtype=X.dtype, device=X.device)
    for i in range(P):
        out[i, :, :, :NXl] = X[:, :, i * NXl:(i + 1) * NXl]
        out[i, :, :, NXl] = X[:, :, ((i + 1) * NXl) % NX]
        out[i, :, :, NXl + 1] = X[:, :, (i * NXl - 1) % NX]
    return out


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

    def all_to_all_start(self, send, recv=None):
        P = self.nranks
        if recv is None:
            return [torch.stack([send[p][r] for p in range(P)]) for r in range(P)]
        for r in range(P):
            for p in range(P):
                recv[r][p].copy_(send[p][r])
        return recv

    def all_to_all_finish(self, handle):
        return handle

    def barrier(self):
        torch.cuda.synchronize()

    def stream_barrier(self):
        """cross-rank barrier in stream order: the slabs of this process share the stream"""

    def stream_barrier_start(self):
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def stream_barrier_finish(self, ev):
        torch.cuda.current_stream().wait_event(ev)

    def map_peers(self, slabs) -> bool:
        """direct peer-memory transport: every slab gets the others' buffers as plain pointers"""
        ptrs = {}
        for s in slabs:
            X, R = C.c_void_p(), C.c_void_p()
            s.ck(s.L.ek_slab_poisson_my_buffers(s.h, C.byref(X), C.byref(R)), "ek_slab_poisson_my_buffers")
            ptrs[s.rank] = (X.value, R.value)
        for s in slabs:
            for r, (X, R) in ptrs.items():
                s.ck(s.L.ek_slab_poisson_set_peer(s.h, r, C.c_void_p(X), C.c_void_p(R)), "ek_slab_poisson_set_peer")
        return True

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

    def all_to_all_start(self, send, recv=None):
        recv = torch.empty_like(send[0]) if recv is None else recv[0]
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

    def stream_barrier(self):
        """cross-rank barrier in stream order (no host synchronisation): a one-element
        all-reduce starts after this rank's preceding kernels and completes only once
        every rank has joined; the kernels issued after it wait for it"""
        if self.nranks > 1:
            if not hasattr(self, "_flag"):
                self._flag = torch.zeros(1, dtype=torch.float32, device="cuda")
            self.dist.all_reduce(self._flag)

    def stream_barrier_start(self):
        """the same barrier, joined on the current stream and awaited (finish) on another one"""
        if self.nranks == 1:
            ev = torch.cuda.Event()
            ev.record()
            return ev
        if not hasattr(self, "_flag"):
            self._flag = torch.zeros(1, dtype=torch.float32, device="cuda")
        return self.dist.all_reduce(self._flag, async_op=True)

    def stream_barrier_finish(self, work):
        if self.nranks == 1:
            torch.cuda.current_stream().wait_event(work)
        else:
            work.wait()

    def map_peers(self, slabs) -> bool:
        """direct peer-memory transport: exchange CUDA IPC handles of the pencil and
        receive buffers; False (on every rank) if any mapping failed"""
        s = slabs[0]
        n = s.L.ek_slab_poisson_ipc_bytes()
        mine = (C.c_ubyte * n)()
        ok = s.L.ek_slab_poisson_ipc_export(s.h, mine) == 0
        t = torch.tensor(list(mine), dtype=torch.uint8, device="cuda")
        every = [torch.empty_like(t) for _ in range(self.nranks)]
        self.dist.all_gather(every, t)
        if ok:
            for r, tr in enumerate(every):
                if r == self.rank:
                    continue
                buf = (C.c_ubyte * n)(*tr.cpu().tolist())
                if s.L.ek_slab_poisson_ipc_import(s.h, r, buf) != 0:
                    ok = False
                    break
        flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
        self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
        return bool(flag.item() > 0.5)

    def max_over_ranks(self, x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def _blocking(comm):
    """blocking forms, shared by both transports"""
    def neighbor_exchange(to_left, to_right, from_left, from_right):
        comm.neighbor_exchange_finish(comm.neighbor_exchange_start(to_left, to_right, from_left, from_right))

    def all_to_all(send, recv=None):
        return comm.all_to_all_finish(comm.all_to_all_start(send, recv))
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
        self._dq_flat = device_view(p.value, (self.NZ * self.NY * self.PX,), self.dev)
        self.dq = self._dq_flat.view(self.NZ, self.NY, self.PX)
        self.sim._ck(self.L.ek_field_ptr(self.h, ek.FIELDS.index("phi"), C.byref(p)), "ek_field_ptr")
        self.phi = device_view(p.value, (self.NZ, self.NY, self.PX), self.dev)
        nh = self.L.ek_halo_doubles(self.h)
        mk = lambda n: torch.empty(n, dtype=torch.float64, device=self.dev)  # noqa: E731
        self.h_to_l, self.h_to_r, self.h_from_l, self.h_from_r = mk(nh), mk(nh), mk(nh), mk(nh)
        np_ = self.NY * self.NZ
        self.p_to_l, self.p_to_r, self.p_from_l, self.p_from_r = mk(np_), mk(np_), mk(np_), mk(np_)
        self.nyh, self.kyl = ky_chunks(self.NY, nranks)
        self.send, self.recv, self.blocks = [], [], []

    def setup_poisson(self, nchunks: int):
        """native distributed Poisson stage: chunk buffers as torch views for the transport"""
        self.ck(self.L.ek_slab_poisson_setup(self.h, int(nchunks)), "ek_slab_poisson_setup")
        K = self.L.ek_slab_poisson_chunks(self.h)
        # c+ - c- is now kept as [y][z][x] rows without ghost columns (EkConst::dq_sy/dq_sz)
        self.dq = self._dq_flat.as_strided((self.NZ, self.NY, self.NX), (self.NX, self.NZ * self.NX, 1))
        self.send, self.recv, self.blocks = [], [], []
        for k in range(K):
            b0, b1, cnt = C.c_int(), C.c_int(), C.c_longlong()
            ps, pr = C.c_void_p(), C.c_void_p()
            self.ck(self.L.ek_slab_poisson_chunk(self.h, k, C.byref(b0), C.byref(b1), C.byref(ps), C.byref(pr),
                                                 C.byref(cnt)), "ek_slab_poisson_chunk")
            n = cnt.value // self.nranks
            self.blocks.append((b0.value, b1.value))
            if cnt.value == 0:
                self.send.append(None)
                self.recv.append(None)
                continue
            self.send.append(device_view(ps.value, (self.nranks, n), self.dev, "<c16"))
            self.recv.append(device_view(pr.value, (self.nranks, n), self.dev, "<c16"))
        return K

    def ck(self, st, what):
        self.sim._ck(st, what)

    # -- Poisson pieces, torch.fft restatement (cross-check of the native stage in the GPU tests) ----
    def poisson_forward_local(self) -> torch.Tensor:
        return y_forward(self.dq[1:1 + self.M, :, :self.NX], self.nranks, self.kyl)

    def poisson_middle(self, recv: torch.Tensor) -> torch.Tensor:
        X = x_forward(recv)
        self.ck(self.L.ek_zsolve_columns(self.h, C.c_void_p(X.data_ptr()), self.rank * self.kyl, self.kyl),
                "ek_zsolve_columns")
        return x_backward(X, self.nranks)

    def poisson_backward_local(self, recv: torch.Tensor):
        self.phi[1:1 + self.M, :, :self.NX] = y_backward(recv, self.NY)
        self.ck(self.L.ek_poisson_finish(self.h, 0), "ek_poisson_finish")


# ---------------------------------------------------------------------------
# a group of slabs driven together (all of them with LocalComm, one with DistComm)
# ---------------------------------------------------------------------------
class SlabGroup:
    def __init__(self, ek, params, comm, device: int = 0, zchunk=None):
        self.ek, self.comm, self.params = ek, comm, params
        self.nranks = comm.nranks
        self.slabs = [Slab(ek, params, r, self.nranks, device, zchunk) for r in comm.local_ranks]
        self.t = 0.0
        self.profile = False          # per-phase CUDA-event timing (development aid)
        self._ev = []
        self.overlap = True           # forward half of the Poisson stage runs behind the LBM launches
        self.overlap_back = True      # ... and the way back behind the next step's first LBM launches
        self.K = 0
        self.side = torch.cuda.Stream(device=self.slabs[0].dev)
        self.halo_stream = torch.cuda.Stream(device=self.slabs[0].dev)
        self.copy_stream = torch.cuda.Stream(device=self.slabs[0].dev)
        self.back_stream = torch.cuda.Stream(device=self.slabs[0].dev)
        self._phi_ready = None        # per-chunk events of the previous step's way back (overlapped steps)
        self.transport = "nccl"       # "nccl": all-to-all of the chunk buffers; "p2p": direct peer-memory writes
        self.set_poisson_chunks(4)

    def set_poisson_chunks(self, nchunks: int):
        """pipeline depth of the Poisson stage (groups of the LBM kernel's z-blocks)"""
        ks = {s.setup_poisson(nchunks) for s in self.slabs}
        assert len(ks) == 1
        self.K = ks.pop()
        if self.transport != "nccl":
            self.set_transport(self.transport)     # the buffers were re-allocated: map them again

    def set_transport(self, transport: str) -> str:
        """"p2p": the re-blocking kernels write straight into the peers' buffers over NVLink
        (CUDA IPC); "dma": the same pushes as strided copies on the copy engines; both fall
        back to "nccl" when the buffers cannot be mapped"""
        if transport in ("p2p", "dma"):
            self.comm.barrier()
            if not self.comm.map_peers(self.slabs):
                transport = "nccl"
            self.comm.barrier()
        for s in self.slabs:
            s.ck(s.L.ek_slab_poisson_set_dma(s.h, int(transport == "dma")), "ek_slab_poisson_set_dma")
        self.transport = transport
        return transport

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

    def phi_halo_exchange_range(self, z0: int, z1: int):
        """ghost columns of phi for the planes [z0, z1) (contiguous slices of the halo buffers)"""
        for s in self.slabs:
            s.ck(s.L.ek_phi_halo_pack_range(s.h, z0, z1, C.c_void_p(s.p_to_l.data_ptr()), C.c_void_p(s.p_to_r.data_ptr())),
                 "ek_phi_halo_pack_range")
        a, b = z0 * self.slabs[0].NY, z1 * self.slabs[0].NY
        self.comm.neighbor_exchange([s.p_to_l[a:b] for s in self.slabs], [s.p_to_r[a:b] for s in self.slabs],
                                    [s.p_from_l[a:b] for s in self.slabs], [s.p_from_r[a:b] for s in self.slabs])
        for s in self.slabs:
            s.ck(s.L.ek_phi_halo_unpack_range(s.h, z0, z1, C.c_void_p(s.p_from_l.data_ptr()),
                                              C.c_void_p(s.p_from_r.data_ptr())), "ek_phi_halo_unpack_range")

    def _chunk_planes(self, k: int):
        """planes of chunk k, wall planes included in the first and last chunk"""
        zc = int(self.slabs[0].sim.counter("zchunk")) if not hasattr(self, "_zc") else self._zc
        self._zc = zc
        b0, b1 = self.slabs[0].blocks[k]
        return b0 * zc, min(b1 * zc, self.slabs[0].NZ)

    def _poisson_tail(self, landed, finish):
        """way back, chunk by chunk on the `back` stream: wait until chunk k has landed, inverse
        y-transform into phi, ghost columns of its planes, then the event that the next LBM
        launches wait for.  The caller decides when the main stream joins (self.join_back())."""
        main = torch.cuda.current_stream()
        self.back_stream.wait_stream(main)
        self._phi_ready = []
        with self._on_stream(self.back_stream):
            # wall planes first (only when something other than the solver wrote phi): the ghost columns
            # of plane 0 travel with chunk 0 and the next step's first LBM launches read them
            for s in self.slabs:
                s.ck(s.L.ek_poisson_finish(s.h, 0), "ek_poisson_finish")
            for k in range(self.K):
                finish(landed[k])
                for s in self.slabs:
                    s.ck(s.L.ek_slab_poisson_backward(s.h, k), "ek_slab_poisson_backward")
                z0, z1 = self._chunk_planes(k)
                self.phi_halo_exchange_range(z0, z1)
                ev = torch.cuda.Event()
                ev.record()
                self._phi_ready.append(ev)

    def join_back(self):
        """the main stream waits for the whole potential (end of a step() call, start-up loop)"""
        torch.cuda.current_stream().wait_stream(self.back_stream)
        self._phi_ready = None

    def _a2a_start(self, k, which):
        bufs = [(s.send[k], s.recv[k]) for s in self.slabs]
        if bufs[0][0] is None:
            return None
        return self.comm.all_to_all_start([b[0] for b in bufs], [b[1] for b in bufs])

    def _a2a_finish(self, hnd):
        if hnd is not None:
            self.comm.all_to_all_finish(hnd)

    def poisson_forward(self, k: int):
        """chunk k: re-blocking + y-transform of my columns, then its transpose starts"""
        for s in self.slabs:
            s.ck(s.L.ek_slab_poisson_forward(s.h, k), "ek_slab_poisson_forward")
        if self.transport != "nccl":
            for s in self.slabs:
                s.ck(s.L.ek_slab_poisson_push_x(s.h, k), "ek_slab_poisson_push_x")
            return None
        return self._a2a_start(k, 0)

    def poisson_rest(self, pending):
        """everything after the forward halves were started: pencils, solve, way back"""
        if self.transport != "nccl":
            return self._poisson_rest_p2p()
        for k, hnd in enumerate(pending):
            self._a2a_finish(hnd)
            for s in self.slabs:
                s.ck(s.L.ek_slab_poisson_gather_x(s.h, k), "ek_slab_poisson_gather_x")
        self._mark("poisson_transpose_1_gather_x")
        for s in self.slabs:
            s.ck(s.L.ek_slab_poisson_solve(s.h), "ek_slab_poisson_solve")
        self._mark("poisson_x_fft_zsolve_x_ifft")
        pending = []
        for k in range(self.K):
            for s in self.slabs:
                s.ck(s.L.ek_slab_poisson_scatter_x(s.h, k), "ek_slab_poisson_scatter_x")
            pending.append(self._a2a_start(k, 1))
        self._mark("poisson_scatter_x")
        self._poisson_tail(pending, self._a2a_finish)

    def poisson_reference(self):
        """the same stage through torch.fft and un-chunked transposes (tests only)"""
        recv = self.comm.all_to_all([s.poisson_forward_local() for s in self.slabs])
        recv = self.comm.all_to_all([s.poisson_middle(r) for s, r in zip(self.slabs, recv)])
        for s, r in zip(self.slabs, recv):
            s.poisson_backward_local(r)
        self.phi_halo_exchange()

    def _poisson_rest_p2p(self):
        self.comm.stream_barrier()            # every rank's rows have landed in my pencils
        self._mark("poisson_transpose_1_barrier")
        for s in self.slabs:
            s.ck(s.L.ek_slab_poisson_solve(s.h), "ek_slab_poisson_solve")
        self._mark("poisson_x_fft_zsolve_x_ifft")
        # pushes of chunk k+1 travel (copy stream) while chunk k is transformed back (main
        # stream); one barrier per chunk tells every rank that its x blocks have landed
        main = torch.cuda.current_stream()
        self.copy_stream.wait_stream(main)
        landed = []
        with self._on_stream(self.copy_stream):
            for k in range(self.K):
                for s in self.slabs:
                    s.ck(s.L.ek_slab_poisson_push_back(s.h, k), "ek_slab_poisson_push_back")
                landed.append(self.comm.stream_barrier_start())
        self._poisson_tail(landed, self.comm.stream_barrier_finish)

    def poisson(self):
        """the distributed fast_Poisson(): dq -> phi (interior, walls, ghost columns).
        The planes are processed in self.K chunks so that the all-to-all of one
        chunk travels while the next chunk is being transformed."""
        pending = [self.poisson_forward(k) for k in range(self.K)]
        self._mark("poisson_y_fft")
        self.poisson_rest(pending)
        self.join_back()
        self._mark("poisson_way_back")

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

    def _on_stream(self, stream):
        """context: torch's current stream and the handles' stream are `stream`"""
        grp = self

        class _Ctx:
            def __enter__(self):
                self.main = torch.cuda.current_stream()
                self.ctx = torch.cuda.stream(stream)
                self.ctx.__enter__()
                for s in grp.slabs:
                    s.ck(s.L.ek_switch_stream(s.h, C.c_void_p(stream.cuda_stream)), "ek_switch_stream")

            def __exit__(self, *exc):
                for s in grp.slabs:
                    s.ck(s.L.ek_switch_stream(s.h, C.c_void_p(self.main.cuda_stream)), "ek_switch_stream")
                self.ctx.__exit__(*exc)
        return _Ctx()

    def lbm_and_forward(self, full: bool):
        """One LBM pass launched chunk by chunk; chunk k's Poisson forward half
        (re-blocking, y-transform, transpose 1) runs on a side stream as soon as
        the launch that produces its planes has finished, behind the launches of
        the later chunks."""
        main = torch.cuda.current_stream()
        pending = []
        for k in range(self.K):
            b0, b1 = self.slabs[0].blocks[k]
            if self._phi_ready is not None:
                # the planes of chunk k take grad(phi) from the chunks k-1 .. k+1 of the previous solve,
                # whose way back may still be running on the back stream
                main.wait_event(self._phi_ready[min(k + 1, self.K - 1)])
            for s in self.slabs:
                s.ck(s.L.ek_stream_collide_save_range(s.h, int(full), b0, b1, int(k == self.K - 1)),
                     "ek_stream_collide_save_range")
            ev = torch.cuda.Event()
            ev.record(main)
            self.side.wait_event(ev)
            with self._on_stream(self.side):
                pending.append(self.poisson_forward(k))
        main.wait_stream(self.side)
        return pending

    def step(self, nsteps: int = 1):
        for i in range(nsteps):
            full = i == nsteps - 1
            parity = self.slabs[0].L.ek_lbm_parity(self.slabs[0].h)
            self._mark("begin")
            if self.overlap and self.K > 1:
                pending = self.lbm_and_forward(full)
                self._mark("lbm_with_poisson_forward")
            else:
                for s in self.slabs:
                    s.sim.stream_collide_save(full)
                self._mark("lbm")
                pending = None
            # the populations travel while the Poisson stage computes (independent data):
            # pack, NCCL send/recv and unpack run on their own stream next to the Poisson kernels
            phase = 0 if parity == 0 else 1
            main = torch.cuda.current_stream()
            self.halo_stream.wait_stream(main)
            with self._on_stream(self.halo_stream):
                hnd = self.halo_exchange_start(phase)
            if pending is None:
                self.poisson()
            else:
                self.poisson_rest(pending)
                if full or not self.overlap_back:
                    self.join_back()
                    self._mark("poisson_way_back")
                # else: the way back of the last chunks runs behind the next step's first LBM launches
            with self._on_stream(self.halo_stream):
                self.halo_exchange_finish(phase, hnd)
            main.wait_stream(self.halo_stream)
            self._mark("population_halo_tail")
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
    if getattr(args, "poisson_chunks", 4) != 4:
        grp.set_poisson_chunks(args.poisson_chunks)
    grp.overlap = not getattr(args, "no_overlap", False)
    transport = grp.set_transport(getattr(args, "transport", "nccl"))
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
    grp.profile = False
    zchunk_used = int(grp.slabs[0].sim.counter("zchunk"))
    # ---- end to end through the C ABI with HOST buffers: every rank uploads its slab of the 11
    # macroscopic arrays from pinned memory, init_equilibrium, K steps, downloads the 11 arrays
    e2e = None
    if not getattr(args, "no_e2e", False):
        try:
            sim = grp.slabs[0].sim
            host = {n: torch.empty(sim.shape, dtype=torch.float64, pin_memory=True).numpy() for n in ek.FIELDS}
            for n in ek.FIELDS:
                sim.field(n, out=host[n])
            comm.barrier()
            t0 = time.perf_counter()
            sim.set_fields(host)
            grp.init_equilibrium()
            grp.step(args.steps)
            for n in ek.FIELDS:
                sim.field(n, out=host[n])
            torch.cuda.synchronize()
            dt = comm.max_over_ranks(time.perf_counter() - t0)
            e2e = {"value": round(cells * args.steps / dt / 1e6, 2), "unit": "MLUPS",
                   "h2d_bytes_per_step": int(11 * cells * 8 / args.steps), "d2h_bytes_per_step": int(11 * cells * 8 / args.steps),
                   "job": f"every rank: upload its slab of 11 fields (pinned host) + init_equilibrium + {args.steps} steps + "
                          "download 11 fields; wall clock, max over ranks", "seconds": round(dt, 4)}
            del host
        except Exception as exc:  # noqa: BLE001  (e.g. the pinned allocation of 11 x 1 GB per rank failed)
            e2e = {"value": None, "unit": "MLUPS", "error": str(exc)[:200]}
    grp.close()
    step_gbs = mlups * 1e6 * B_ALG_STEP / 1e9
    return {"metric": "coupled_step_mlups", "value": round(mlups, 2), "unit": "MLUPS", "n_gpus": comm.nranks,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak" if w.get("weak") else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "grid": [NX, NY, NZ], "stream_mode": "aa", "zchunk": zchunk_used,
                       "parallelism": f"x-slabs x{comm.nranks}: NCCL halo send/recv + Poisson transposes by "
                                      + {"nccl": "NCCL all-to-all, ", "p2p": "direct peer-memory writes (CUDA IPC, kernel), ",
                                         "dma": "direct peer-memory copies (CUDA IPC, copy engines), "}[transport] +
                                      f"{grp.K} z-chunks, forward half overlapped with the LBM launches",
                       "cells_per_gpu": cells // comm.nranks, "init": "reference start-up (PB iterations) %.2f s" % init_s,
                       "l2": "per-GPU working set >> 126 MB L2", "phase_ms_rank0": phases},
            "roofline": {"bound": "hbm", "achieved": round(step_gbs / comm.nranks, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(step_gbs / comm.nranks / peak, 4), "peak_source": peak_src,
                         "note": "whole coupled step per GPU at 1760 B/cell (kernel split is reported at N=1)",
                         "traffic": None},
            "cpu_baseline": None,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "hbm_gbs_step": round(step_gbs, 1)}
