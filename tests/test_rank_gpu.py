"""The native one-process-per-GPU driver (csrc/ek_rank.cu, ek_rank_* of the C ABI).  On the 1-GPU test
box it runs as a single rank (no NCCL traffic, same stream pipeline); with two or more GPUs the 2-rank
NCCL run of tools/rank_check.py is launched under torchrun as well."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch  # noqa: F401  (before anything dlopens libnccl.so.2: PyTorch needs its own copy)

from tests import util
from tests.test_oracle_cpu import check
from tests.test_parity_gpu import oracle_run, product_run, synthetic_init

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ek():
    return util.ek_module()


@pytest.mark.parametrize("chunks,NX", [(1, 64), (3, 64), (3, 96), (2, 128)])
def test_single_rank_matches_the_single_domain_run_and_the_oracle(ek, chunks, NX):
    """NX >= 96: three or more x-tiles per row, so the LBM pass is launched boundary tiles first and the
    halo exchange runs under the interior launches"""
    over = dict(NX=NX, NY=6, NZ=21, exf=1.0e6, uw=1.0e-4, voltage2=-3.0e-3)
    init = synthetic_init(over)
    want, _ = product_run(ek, over, init, 7, ek.STREAM_AA)
    ref, _ = oracle_run(over, init, 7)
    res = []
    for overlap, bfirst in ((True, True), (False, False), (True, False)):
        rs = ek.RankSimulation(ek.default_params(**over), 0, 0, 1, None, poisson_chunks=chunks)
        assert rs.chunks() == chunks
        rs.set_pipeline(overlap, overlap, boundary_first=bfirst)
        rs.set_fields(init)
        rs.init_equilibrium()
        rs.step(4)
        rs.step(3)
        res.append(rs.fields())
        assert rs.counter("steps") == 7
        rs.close()
    check(util.field_errors(res[0], want))
    check(util.field_errors(res[0], ref))
    for other in res[1:]:
        for k in util.FIELDS:
            assert np.array_equal(res[0][k], other[k]), k      # pipelined / boundary-first == sequential, bit for bit


def test_single_rank_startup_timed_steps_and_profile(ek):
    over = dict(NX=32, NY=4, NZ=13, pb_iters=40)
    sim = ek.Simulation(ek.default_params(**over))
    sim.init()
    sim.step(3)
    want = sim.fields()
    sim.close()
    rs = ek.RankSimulation(ek.default_params(**over), 0, 0, 1, None)
    with pytest.raises(ek.EkError):
        rs.step(1)                                  # before init_equilibrium
    rs.init()
    ms = rs.step_timed(3)
    assert ms > 0.0
    check(util.field_errors(rs.fields(), want))
    for sequential in (False, True):
        ph = rs.profile(2, sequential)
        assert "x_fft_zsolve_x_ifft" in ph and all(v >= 0.0 for v in ph.values())
    assert rs.counter("steps") == 3 + 4
    # a slab handle of a multi-rank domain refuses the single-domain entry points
    h = ek.Simulation(ek.default_params(**over), slab=(0, 2))
    with pytest.raises(ek.EkError):
        h.init()
    with pytest.raises(ek.EkError):
        h.set_option("stream_mode", ek.STREAM_PUSH)
    h.close()
    rs.close()


def test_nccl_can_be_loaded(ek):
    L = ek.load_library()
    assert L.ek_rank_nccl_id_bytes() == 3 * 128
    assert L.ek_rank_nccl_version() >= 20000


def test_two_ranks_over_nccl_match_the_single_domain_run(ek):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (tools/rank_check.py under torchrun; run by the builder with gpurun --gpus 2)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(util.ROOT, "tools", "rank_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    rep = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert rep["ok"] and rep["world"] == 2
