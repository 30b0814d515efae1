"""Host-side logic of the multi-GPU path on CPU: partitioning, the transports
(torch.distributed with gloo, world_size 2) and the transposes of the
distributed Poisson stage, with the device z-solve replaced by a dense solve.
The result is checked against the oracle's fast_Poisson."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ek_oracle as eo
from tests import util


def slab_mod():
    import importlib
    util.ek_module()
    return importlib.import_module("ek-pnp-3d_b200.slab")


def test_partition_and_ky_chunks():
    slab = slab_mod()
    assert slab.partition(1024, 8)[3] == (384, 512)
    with pytest.raises(ValueError):
        slab.partition(50, 4)
    nyh, kyl = slab.ky_chunks(256, 8)
    assert (nyh, kyl) == (129, 17) and kyl * 8 >= nyh
    assert slab.ky_chunks(8, 2) == (5, 3)


def zsolve_dense(X, p, ky0, NXg, NY):
    """what ek_zsolve_columns does, as dense solves (numpy, tiny grids only)"""
    M, kyl, _ = X.shape
    out = np.zeros_like(X)
    nxy = NXg * NY
    for iy in range(kyl):
        ky = ky0 + iy
        J = (ky if ky <= NY // 2 else ky - NY) * 2 * np.pi / p.Ly
        for ix in range(NXg):
            I = (ix if ix <= NXg // 2 else ix - NXg) * 2 * np.pi / p.Lx
            A = (np.diag(np.full(M, -(2.0 + (I * I + J * J) * p.dz ** 2))) + np.diag(np.ones(M - 1), 1)
                 + np.diag(np.ones(M - 1), -1))
            d = -(p.convertCtoCharge / p.eps) * p.dz ** 2 * X[:, iy, ix]
            if ky == 0 and ix == 0:
                d = d.copy()
                d[0] += -p.voltage * nxy
                d[-1] += -p.voltage2 * nxy
            out[:, iy, ix] = np.linalg.solve(A, d) / nxy
    return out


def run_distributed_poisson(slab, comm, p, dq_local, ranks, ghosts=False):
    """dq_local: {rank: (NZ, NY, NXl)} -> {rank: phi interior (M, NY, NXl)}; ghosts: (M, NY, NXl + 2), the last
    two columns being the right / left ghost column of phi (they travel inside the second transpose)"""
    P = comm.nranks
    nyh, kyl = slab.ky_chunks(p.NY, P)
    send = [slab.y_forward(torch.from_numpy(dq_local[r][1:-1].copy()), P, kyl) for r in ranks]
    recv = comm.all_to_all(send)
    back = []
    for r, rc in zip(ranks, recv):
        X = slab.x_forward(rc)
        X = torch.from_numpy(zsolve_dense(X.numpy(), p, r * kyl, p.NX, p.NY))
        back.append(slab.x_backward_ghost(X, P) if ghosts else slab.x_backward(X, P))
    recv = comm.all_to_all(back)
    return {r: slab.y_backward(rc, p.NY).numpy() for r, rc in zip(ranks, recv)}


def oracle_case():
    p = eo.default_params(NX=12, NY=6, NZ=8, voltage2=-2.0e-3)
    o = eo.Oracle(p)
    o.set_poisson_dc(0)
    rng = np.random.default_rng(3)
    f = o.fields()
    f["charge"] = 0.01 * (1 + 0.1 * rng.standard_normal(o.shape))
    f["chargen"] = 0.01 * (1 + 0.1 * rng.standard_normal(o.shape))
    o.set_fields(f)
    o.fast_poisson()
    return p, f["charge"] - f["chargen"], o.field("phi").copy()


def test_local_transport_fills_preallocated_receive_buffers():
    slab = slab_mod()
    P = 3
    comm = slab.LocalComm(P)
    send = [torch.stack([torch.full((2,), 10.0 * r + i) for i in range(P)]) for r in range(P)]
    recv = [torch.zeros_like(s) for s in send]
    out = comm.all_to_all(send, recv)
    for r in range(P):
        assert out[r] is recv[r]
        for i in range(P):
            assert bool((recv[r][i] == 10.0 * i + r).all())


@pytest.mark.parametrize("P", [1, 2, 3, 4])
def test_local_transport_matches_the_oracle(P):
    slab = slab_mod()
    p, dq, phi = oracle_case()
    parts = slab.partition(p.NX, P)
    comm = slab.LocalComm(P)
    got = run_distributed_poisson(slab, comm, p, {r: dq[:, :, a:b] for r, (a, b) in enumerate(parts)}, list(range(P)))
    full = np.concatenate([got[r] for r in range(P)], axis=2)
    assert np.abs(full - phi[1:-1]).max() <= 1e-12 * np.abs(phi).max()


@pytest.mark.parametrize("P", [1, 2, 3])
def test_ghost_columns_travel_inside_the_second_transpose(P):
    """the way back of the native rank driver: rows two columns wider, the ghost columns of phi come out of
    the inverse y-transform -- they must be the neighbours' edge columns (periodic in x)"""
    slab = slab_mod()
    p, dq, phi = oracle_case()
    parts = slab.partition(p.NX, P)
    comm = slab.LocalComm(P)
    got = run_distributed_poisson(slab, comm, p, {r: dq[:, :, a:b] for r, (a, b) in enumerate(parts)}, list(range(P)),
                                  ghosts=True)
    scale = np.abs(phi).max()
    for r, (a, b) in enumerate(parts):
        assert got[r].shape[2] == (b - a) + 2
        assert np.abs(got[r][:, :, :b - a] - phi[1:-1, :, a:b]).max() <= 1e-12 * scale
        assert np.abs(got[r][:, :, b - a] - phi[1:-1, :, b % p.NX]).max() <= 1e-12 * scale          # right ghost
        assert np.abs(got[r][:, :, b - a + 1] - phi[1:-1, :, (a - 1) % p.NX]).max() <= 1e-12 * scale  # left ghost


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        slab = slab_mod()
        comm = slab.DistComm(dist)
        # ring exchange: what my neighbours sent must arrive on the right side
        to_l, to_r = torch.full((4,), 10.0 * rank + 1), torch.full((4,), 10.0 * rank + 2)
        fl, fr = torch.zeros(4), torch.zeros(4)
        comm.neighbor_exchange([to_l], [to_r], [fl], [fr])
        left, right = (rank - 1) % world, (rank + 1) % world
        ok = bool((fl == 10.0 * left + 2).all() and (fr == 10.0 * right + 1).all())
        p, dq, phi = oracle_case()
        a, b = slab.partition(p.NX, world)[rank]
        got = run_distributed_poisson(slab, comm, p, {rank: dq[:, :, a:b]}, [rank])[rank]
        err = float(np.abs(got - phi[1:-1, :, a:b]).max() / np.abs(phi).max())
        gg = run_distributed_poisson(slab, comm, p, {rank: dq[:, :, a:b]}, [rank], ghosts=True)[rank]
        err = max(err, float(np.abs(gg[:, :, b - a] - phi[1:-1, :, b % p.NX]).max() / np.abs(phi).max()),
                  float(np.abs(gg[:, :, b - a + 1] - phi[1:-1, :, (a - 1) % p.NX]).max() / np.abs(phi).max()))
        m = comm.max_over_ranks(float(rank))
        # all-to-all into preallocated chunk buffers (what the native Poisson stage hands to the transport):
        # part i of my send buffer must arrive as part `rank` of rank i's receive buffer
        send = torch.stack([torch.full((3,), 100.0 * rank + i, dtype=torch.complex128) for i in range(world)])
        recv = torch.zeros_like(send)
        comm.all_to_all_finish(comm.all_to_all_start([send], [recv]))
        ok = ok and all(bool((recv[i] == 100.0 * i + rank).all()) for i in range(world))
        q.put((rank, ok, err, m))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, ok, err, m in res:
        assert ok, f"rank {rank}: neighbour exchange delivered the wrong buffers"
        assert err <= 1e-12, (rank, err)
        assert m == 1.0


def plan_chunks(nblocks, nchunks, sizes=None):
    import ctypes as C
    lib = util.ek_module().load_library()
    fn = lib.ek_slab_poisson_plan_chunks
    fn.restype = C.c_int
    fn.argtypes = [C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_int)]
    bounds = (C.c_int * 17)()
    K = fn(nblocks, nchunks, None if sizes is None else sizes.encode(), bounds)
    return K, list(bounds[:K + 1])


def test_chunk_plan_of_the_pipelined_poisson_stage():
    """ek_slab_poisson_plan_chunks (host arithmetic of csrc/ek_slab_poisson.cu): every plan covers the z-blocks
    exactly once with non-empty chunks; the automatic plan is the measured 1:2:3:4:3:2:1 one on large grids."""
    # C5 per rank: 256 planes in z-blocks of 8 -> 32 blocks, automatic
    K, b = plan_chunks(32, 0)
    assert K == 7 and b == [0, 2, 6, 12, 20, 26, 30, 32]
    assert [b[i + 1] - b[i] for i in range(K)] == [2, 4, 6, 8, 6, 4, 2]
    # below 16 z-blocks: up to four equal chunks; tiny grids: one chunk per block
    assert plan_chunks(15, 0) == (4, [0, 3, 7, 11, 15])
    assert plan_chunks(3, 0) == (3, [0, 1, 2, 3])
    assert plan_chunks(1, 0) == (1, [0, 1])
    # explicit counts: equal chunks, never more than z-blocks or the compiled-in maximum
    assert plan_chunks(32, 4) == (4, [0, 8, 16, 24, 32])
    assert plan_chunks(5, 8) == (5, [0, 1, 2, 3, 4, 5])
    assert plan_chunks(64, 40)[0] == 16
    # explicit sizes win when they are valid, and are ignored otherwise
    assert plan_chunks(16, 0, "2,3,4,4,2,1") == (6, [0, 2, 5, 9, 13, 15, 16])
    assert plan_chunks(16, 2, "2,3,4,4,2") == (2, [0, 8, 16])          # sum != nblocks
    assert plan_chunks(16, 2, "8,0,8") == (2, [0, 8, 16])              # empty chunk
    assert plan_chunks(17, 2, ",".join(["1"] * 17)) == (2, [0, 8, 17])  # more chunks than the maximum
    # every automatic / counted plan: strictly increasing, complete
    for nblocks in range(1, 200):
        for nchunks in (0, 1, 2, 3, 4, 7, 16, 50):
            K, b = plan_chunks(nblocks, nchunks)
            assert 1 <= K <= 16 and b[0] == 0 and b[-1] == nblocks
            assert all(b[i + 1] > b[i] for i in range(K)), (nblocks, nchunks, b)
    assert plan_chunks(0, 0)[0] == 0
