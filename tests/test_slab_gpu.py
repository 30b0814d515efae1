"""The multi-slab algorithm (halo phases A/B, phi halos, distributed Poisson,
host-driven start-up) checked on ONE GPU: all slabs of the group live in this
process and the collectives are copies (slab.LocalComm).  Reference = the
single-domain run of the same library and the oracle."""
import importlib

import numpy as np
import pytest

from oracle import ek_oracle as eo
from tests import util
from tests.test_oracle_cpu import check
from tests.test_parity_gpu import oracle_run, product_run, synthetic_init

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ek():
    return util.ek_module()


@pytest.fixture(scope="module")
def slab(ek):
    return importlib.import_module("ek-pnp-3d_b200.slab")


@pytest.mark.parametrize("P", [1, 2, 4])
@pytest.mark.parametrize("steps", [1, 2, 5])
def test_slabs_match_the_single_domain_run(ek, slab, P, steps):
    over = dict(NX=64, NY=6, NZ=13)
    init = synthetic_init(over)
    want, wantP = product_run(ek, over, init, steps, ek.STREAM_AA, pops=True)
    grp = slab.SlabGroup(ek, ek.default_params(**over), slab.LocalComm(P))
    grp.set_fields(init)
    grp.init_equilibrium()
    grp.step(steps)
    got = grp.gather_fields()
    check(util.field_errors(got, want))
    for s in range(4):
        a = grp.gather_populations(s)
        assert np.abs(a - wantP[s]).max() <= 1e-13 * np.abs(wantP[s]).max(), s
    grp.close()


def test_slabs_match_the_oracle_with_walls_moving(ek, slab):
    over = dict(NX=48, NY=4, NZ=11, uw=1.0e-4, exf=2.0e6, voltage2=-2.5e-3)
    init = synthetic_init(over)
    want, _ = oracle_run(over, init, 6)
    grp = slab.SlabGroup(ek, ek.default_params(**over), slab.LocalComm(3))
    grp.set_fields(init)
    grp.init_equilibrium()
    grp.step(3)
    grp.step(3)
    check(util.field_errors(grp.gather_fields(), want))
    grp.close()


def test_slab_startup_matches_the_single_domain_startup(ek, slab):
    over = dict(NX=32, NY=4, NZ=11, pb_iters=40)
    sim = ek.Simulation(ek.default_params(**over))
    sim.initialization()
    want = sim.fields()
    sim.close()
    grp = slab.SlabGroup(ek, ek.default_params(**over), slab.LocalComm(2))
    grp.initialization()
    check(util.field_errors(grp.gather_fields(), want))
    grp.init_equilibrium()
    grp.step(2)
    o = eo.Oracle(eo.default_params(**over))
    o.set_poisson_dc(0)
    o.initialization()
    o.init_equilibrium()
    o.step(2)
    check(util.field_errors(grp.gather_fields(), o.fields()))
    grp.close()


def test_single_handle_refuses_the_distributed_solve(ek):
    sim = ek.Simulation(ek.default_params(NX=32, NY=2, NZ=7), slab=(0, 2))
    sim.set_fields({"rho": np.full(sim.shape, 1000.0)})
    sim.init_equilibrium()
    with pytest.raises(ek.EkError):
        sim.fast_Poisson()
    sim.close()


@pytest.mark.parametrize("P,chunks,zchunk", [(2, 1, 4), (2, 3, 4), (4, 4, 3)])
def test_native_distributed_poisson_matches_the_torch_fft_restatement(ek, slab, P, chunks, zchunk):
    """ek_slab_poisson.cu (cuFFT + hand-written re-blocking, chunked along z) against
    the un-chunked torch.fft restatement of the same stage, on the same c+ - c-"""
    over = dict(NX=16 * P, NY=6, NZ=14, voltage2=-3.0e-3)
    init = synthetic_init(over)
    res = []
    for native in (True, False):
        grp = slab.SlabGroup(ek, ek.default_params(**over), slab.LocalComm(P), zchunk=zchunk)
        grp.set_poisson_chunks(chunks)
        assert grp.K == min(chunks, -(-14 // zchunk))
        grp.set_fields(init)
        grp.init_equilibrium()          # fills c+ - c-
        if native:
            grp.poisson()
        else:
            grp.poisson_reference()
        res.append(grp.gather_fields()["phi"])
        grp.close()
    scale = np.abs(res[1]).max()
    assert np.abs(res[0] - res[1]).max() <= 1e-13 * scale


def test_overlapped_and_plain_slab_steps_are_bitwise_identical(ek, slab):
    over = dict(NX=64, NY=4, NZ=13, exf=1.0e6)
    init = synthetic_init(over)
    res = []
    for overlap in (True, False):
        grp = slab.SlabGroup(ek, ek.default_params(**over), slab.LocalComm(2), zchunk=4)
        grp.overlap = overlap
        grp.set_fields(init)
        grp.init_equilibrium()
        grp.step(2)
        grp.step(3)
        res.append(grp.gather_fields())
        grp.close()
    for k in util.FIELDS:
        for other in res[1:]:
            assert np.array_equal(res[0][k], other[k]), k


def test_peer_memory_transport_is_bitwise_identical_to_the_all_to_all(ek, slab):
    """push_x / push_back (re-blocking kernels writing straight into the peers' buffers)
    against all-to-all + local re-blocking: same numbers, same order, so bit for bit"""
    over = dict(NX=96, NY=6, NZ=13, exf=1.0e6, voltage2=-3.0e-3)
    init = synthetic_init(over)
    res = []
    for transport in ("nccl", "p2p", "dma"):
        grp = slab.SlabGroup(ek, ek.default_params(**over), slab.LocalComm(3), zchunk=4)
        assert grp.set_transport(transport) == transport
        grp.set_fields(init)
        grp.init_equilibrium()
        grp.step(2)
        grp.step(3)
        res.append(grp.gather_fields())
        grp.close()
    for k in util.FIELDS:
        for other in res[1:]:
            assert np.array_equal(res[0][k], other[k]), k


@pytest.mark.parametrize("P", [1, 2, 4])
def test_native_multi_driver_matches_the_python_slab_driver(ek, slab, P):
    """ek_multi.cu (one process, peer copies, event barriers) against slab.py (LocalComm) and the oracle:
    same kernels in the same order, so the fields agree bit for bit"""
    over = dict(NX=32 * P, NY=6, NZ=13, exf=1.0e6, voltage2=-3.0e-3)
    init = synthetic_init(over)
    grp = slab.SlabGroup(ek, ek.default_params(**over), slab.LocalComm(P))   # same (automatic) chunking as ek_multi
    grp.set_fields(init)
    grp.init_equilibrium()
    grp.step(2)
    grp.step(3)
    want = grp.gather_fields()
    grp.close()
    m = ek.MultiSimulation(ek.default_params(**over), [0] * P)
    assert m.L.ek_multi_slabs(m.h) == P and m.L.ek_multi_slab(m.h, P) is None
    m.set_fields(init)
    m.init_equilibrium()
    m.step(2)
    m.step(3)
    got = m.fields()
    m.close()
    ref, _ = oracle_run(over, init, 5)
    check(util.field_errors(got, ref))
    for k in util.FIELDS:
        assert np.array_equal(got[k], want[k]), k
    # without the stream pipeline: same numbers
    m = ek.MultiSimulation(ek.default_params(**over), [0] * P)
    assert m.L.ek_multi_set_pipeline(m.h, 0) == 0
    m.set_fields(init)
    m.init_equilibrium()
    m.step(2)
    m.step(3)
    plain = m.fields()
    m.close()
    for k in util.FIELDS:
        assert np.array_equal(plain[k], want[k]), k


def test_native_multi_driver_startup(ek):
    over = dict(NX=32, NY=4, NZ=11, pb_iters=40)
    sim = ek.Simulation(ek.default_params(**over))
    sim.init()
    sim.step(2)
    want = sim.fields()
    sim.close()
    m = ek.MultiSimulation(ek.default_params(**over), [0, 0])
    m.init()
    m.step(2)
    check(util.field_errors(m.fields(), want))
    m.close()


def test_checkpoints_are_interchangeable_between_single_and_multi_gpu_runs(ek, tmp_path):
    """a checkpoint written by a single-domain run continues bit for bit on three slabs and back"""
    over = dict(NX=96, NY=5, NZ=13, exf=1.0e6, uw=1.0e-4)
    init = synthetic_init(over)
    sim = ek.Simulation(ek.default_params(**over))
    sim.set_fields(init)
    sim.init_equilibrium()
    sim.step(3)
    a = str(tmp_path / "single.ekc")
    sim.checkpoint_save(a, 3.0)
    sim.step(4)
    want = sim.fields()
    sim.close()
    m = ek.MultiSimulation(ek.default_params(**over), [0, 0, 0])
    assert m.checkpoint_load(a) == 3.0
    m.step(2)                          # now at A-A parity 0 again after two steps ...
    m.step(1)                          # ... and at parity 1 when saved
    b = str(tmp_path / "multi.ekc")
    m.checkpoint_save(b, 6.0)
    m.step(1)
    got = m.fields()
    m.close()
    check(util.field_errors(got, want))          # slabs vs single domain: distributed solve, not bitwise
    sim = ek.Simulation(ek.default_params(**over))
    assert sim.checkpoint_load(b) == 6.0
    sim.step(1)
    back = sim.fields()
    sim.close()
    check(util.field_errors(back, want))


def test_multi_diagnostics_and_step_counter(ek):
    """ek_multi_wall_current / ek_multi_max_uz against the single-domain diagnostics, with the slabs on
    DISTINCT devices when the box has them (the per-slab reductions select their own device), and the
    step counter that ek_multi_checkpoint_save writes"""
    import torch
    over = dict(NX=64, NY=6, NZ=13, exf=1.0e6, voltage2=-3.0e-3)
    init = synthetic_init(over)
    sim = ek.Simulation(ek.default_params(**over))
    sim.set_fields(init)
    sim.init_equilibrium()
    sim.step(5)
    want_i, want_u = sim.current(), sim.max_uz()
    sim.close()
    ndev = torch.cuda.device_count()
    m = ek.MultiSimulation(ek.default_params(**over), [0, 1 % ndev])
    m.set_fields(init)
    m.init_equilibrium()
    m.step(5)
    assert abs(m.current() - want_i) <= 1e-10 * abs(want_i)
    assert abs(m.max_uz() - want_u) <= 1e-12 * abs(want_u) + 16 * 5.7e-15
    v = ek.C.c_double()
    assert m.L.ek_get_counter(ek.C.c_void_p(m.L.ek_multi_slab(m.h, 1)), b"steps", ek.C.byref(v)) == 0 and v.value == 5
    m.close()


def test_c4_shaped_slabs_match_the_single_domain_run(ek):
    """config C4's decomposition on one device: NX = 1024 split into 8 x-slabs of 128 columns (four x-tiles +
    the ghost tile, the row-stride-immediate kernel instantiation for slabs), small NY / NZ, pressure- and
    electro-driven, 3-D perturbed start; all 11 fields against the single-domain run"""
    over = dict(NX=1024, NY=8, NZ=37, exf=2.0e6, chargeinf=0.002)
    init = synthetic_init(over)
    want, _ = product_run(ek, over, init, 6, ek.STREAM_AA)
    m = ek.MultiSimulation(ek.default_params(**over), [0] * 8)
    m.set_fields(init)
    m.init_equilibrium()
    m.step(4)
    m.step(2)
    got = m.fields()
    m.close()
    check(util.field_errors(got, want))
    # and through the per-rank driver's pipeline pieces on one rank (ghost columns inside the transpose)
    rs = ek.RankSimulation(ek.default_params(**over), 0, 0, 1, None)
    rs.set_fields(init)
    rs.init_equilibrium()
    rs.step(6)
    check(util.field_errors(rs.fields(), want))
    rs.close()
