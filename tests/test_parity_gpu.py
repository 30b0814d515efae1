"""Parity of the CUDA path (through the C ABI) with the reference's golden
vectors, with the CPU restatement, and -- when its binaries travelled to the
box -- with the reference's own CUDA build run live.

Tolerances are max|a-b|/max|b| per field group (components of a vector share
the scale); see DESIGN.md "Tolerances" for why the velocity has its own."""
import json
import os

import numpy as np
import pytest

from oracle import ek_oracle as eo
from tests import util
from tests.test_oracle_cpu import K_U, TOL, check, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ek():
    return util.ek_module()


def _record(name, err):
    """measured errors of the live/golden comparisons, for DESIGN.md section 6 (gpurun_out/parity_r02.json)"""
    out = os.path.join(util.ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        path = os.path.join(out, "parity_r02.json")
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[name] = {k: float(v) for k, v in err.items()}
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass


def product_run(ek, over, init, steps, mode, zchunk=None, dc=None, dc_mode=0, pops=False, every=None):
    sim = ek.Simulation(ek.default_params(**over), stream_mode=mode, zchunk=zchunk)
    sim.set_fields(init)
    sim.init_equilibrium()
    if dc is not None:
        for n in range(steps):
            sim.set_poisson_dc(ek.DC_PRESCRIBED, float(dc[n]))
            sim.step(1)
    else:
        sim.set_poisson_dc(dc_mode)
        sim.step(steps)
    f = sim.fields()
    P = np.stack([sim.populations(s) for s in range(4)]) if pops else None
    sim.close()
    return f, P


def oracle_run(over, init, steps, dc_mode=0, pops=False):
    o = eo.Oracle(eo.default_params(**over))
    o.set_fields(init)
    o.init_equilibrium()
    o.set_poisson_dc(dc_mode)
    o.step(steps)
    f = o.fields()
    P = np.stack([o.populations(s) for s in range(4)]) if pops else None
    o.close()
    return f, P


def synthetic_init(over, amp=0.05, pb_iters=30):
    """Reference-style start-up (shortened PB loop) plus the SURVEY 8(d) perturbation."""
    o = eo.Oracle(eo.default_params(**dict(over, pb_iters=pb_iters)))
    o.set_poisson_dc(0)
    o.initialization()
    f = eo.perturb_fields(o.fields(), amp) if amp else o.fields()
    o.close()
    return f


# ---------------------------------------------------------------------------
# against the reference's golden vectors
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("fixture", ["g1_p05_s50.npz", "g3_p05_s50.npz", "g4_p05_s50.npz"])
@pytest.mark.parametrize("mode", [0, 1])
def test_matches_the_reference_golden_run(ek, fixture, mode):
    z, meta = load_golden(fixture)
    init = {k: z[f"init_{k}"] for k in util.FIELDS}
    got, P = product_run(ek, meta["overrides"], init, meta["steps"], mode, dc=z["dc"], pops=True)
    err = util.field_errors(got, {k: z[f"final_{k}"] for k in util.FIELDS})
    _record(f"golden_{fixture}_mode{mode}", err)
    check(err)
    ref = z["final_fluid_pops"]
    assert np.abs(P[0] - ref).max() <= 1e-13 * np.abs(ref).max()


def test_startup_matches_the_reference(ek):
    """ek_init_fields vs the reference's initialization() (501 PB iterations)."""
    z, meta = load_golden("g4_startup.npz")
    sim = ek.Simulation(ek.default_params(**meta["overrides"]))
    sim.initialization()
    got = sim.fields()
    sim.close()
    check(util.field_errors(got, {k: z[f"init_{k}"] for k in util.FIELDS}))


def test_shipped_case_matches_the_reference(ek):
    """Config C1 (50x8x51, 1000 steps, as shipped): NX is not a multiple of the
    32-cell tile, so the masked lanes are exercised."""
    z, meta = load_golden("c1_shipped_s1000.npz")
    shape = (meta["overrides"]["NZ"], meta["overrides"]["NY"], meta["overrides"]["NX"])
    init = {k: np.ascontiguousarray(np.broadcast_to(z[f"init_{k}_zprofile"][:, None, None], shape)) for k in util.FIELDS}
    got, _ = product_run(ek, meta["overrides"], init, meta["steps"], 0, dc=z["dc"])
    for k in ("rho", "charge", "chargen", "phi", "T", "Ez"):
        want = z[f"final_{k}_zprofile"]
        assert np.abs(got[k][:, 0, 0] - want).max() <= 1e-11 * np.abs(want).max(), k
        assert np.abs(got[k] - got[k][:, :1, :1]).max() <= 1e-12 * np.abs(want).max(), k
    want = z["final_ux_zprofile"]
    ulp = util.u_ulp({"rho": z["final_rho_zprofile"]})
    du = np.abs(got["ux"][:, 0, 0] - want).max()
    _record("c1_shipped_1000_steps", {"u_ulps": du / ulp, "u_rel": du / np.abs(want).max()})
    assert du <= 1e-12 * np.abs(want).max() + K_U * ulp, (du / ulp, "ulps")


# ---------------------------------------------------------------------------
# against the CPU restatement, same DC convention on both sides
# ---------------------------------------------------------------------------
CASES = [
    dict(NX=16, NY=8, NZ=13),
    dict(NX=50, NY=3, NZ=9),                                   # ragged x tile, odd NY
    dict(NX=33, NY=1, NZ=5),                                   # minimum NZ, NY = 1, 1-lane tail tile
    dict(NX=24, NY=8, NZ=11, uw=1.0e-4, exf=2.0e6, voltage2=-2.5e-3),  # moving wall, body force
    dict(NX=20, NY=6, NZ=12, TH=0.0),                          # isothermal (T == 0 exactly)
    dict(NX=64, NY=4, NZ=10, Ext=0.0, Ra=5.0),
]


@pytest.mark.parametrize("over", CASES, ids=lambda o: "x".join(str(o[k]) for k in ("NX", "NY", "NZ")))
@pytest.mark.parametrize("steps", [1, 2, 7])
def test_matches_the_oracle(ek, over, steps):
    init = synthetic_init(over)
    want, wantP = oracle_run(over, init, steps, pops=True)
    for mode in (ek.STREAM_AA, ek.STREAM_PUSH):
        got, gotP = product_run(ek, over, init, steps, mode, pops=True)
        err = util.field_errors(got, want)
        if over.get("TH") == 0.0:
            assert np.all(got["T"] == 0.0)
        check(err)
        for s, e in enumerate(util.pop_errors(gotP, wantP)):
            assert e <= 1e-13, (s, e)


def test_stream_modes_and_zchunks_are_bitwise_identical(ek):
    over = dict(NX=40, NY=5, NZ=21)
    init = synthetic_init(over)
    base, baseP = product_run(ek, over, init, 5, ek.STREAM_AA, zchunk=8, pops=True)
    for mode, zc in ((ek.STREAM_AA, 2), (ek.STREAM_AA, 5), (ek.STREAM_AA, 64), (ek.STREAM_PUSH, 3), (ek.STREAM_PUSH, 8)):
        got, gotP = product_run(ek, over, init, 5, mode, zchunk=zc, pops=True)
        for k in util.FIELDS:
            assert np.array_equal(got[k], base[k]), (mode, zc, k)
        assert np.array_equal(gotP, baseP), (mode, zc)


def test_stage_by_stage(ek):
    """stream_collide_save() and fast_Poisson() separately, as the reference's
    loop calls them (main.cu:192,198)."""
    over = dict(NX=32, NY=4, NZ=13)
    init = synthetic_init(over)
    o = eo.Oracle(eo.default_params(**over))
    o.set_poisson_dc(0)
    o.set_fields(init)
    o.init_equilibrium()
    sim = ek.Simulation(ek.default_params(**over))
    sim.set_fields(init)
    sim.init_equilibrium()
    for it in range(3):
        o.stream_collide_save()
        sim.stream_collide_save(True)
        a, b = sim.fields(), o.fields()
        for k in ("rho", "charge", "chargen", "T"):
            assert np.abs(a[k] - b[k]).max() <= 1e-13 * np.abs(b[k]).max(), (it, k)
        assert np.abs(a["ux"] - b["ux"]).max() <= 1e-12 * np.abs(b["ux"]).max() + K_U * util.u_ulp(b)
        # phi and E are untouched by the LBM pass
        assert np.array_equal(a["phi"], sim.field("phi"))
        o.fast_poisson()
        sim.fast_Poisson(True)
        a, b = sim.fields(), o.fields()
        for k in ("phi", "Ez"):
            assert np.abs(a[k] - b[k]).max() <= 1e-12 * np.abs(b[k]).max(), (it, k)
    sim.close()


def test_literal_dc_mode_is_a_constant_interior_shift(ek):
    over = dict(NX=20, NY=6, NZ=12)
    init = synthetic_init(over)
    phis = {}
    for mode in (ek.DC_ZERO, ek.DC_LITERAL):
        sim = ek.Simulation(ek.default_params(**over), xcheck=True)   # the literal transform is cross-check only
        sim.set_fields(init)
        sim.init_equilibrium()
        sim.set_poisson_dc(mode)
        sim.step(1)
        phis[mode] = sim.field("phi")
        sim.close()
    d = (phis[ek.DC_LITERAL] - phis[ek.DC_ZERO])
    assert np.all(d[0] == 0) and np.all(d[-1] == 0)
    assert np.ptp(d[1:-1]) <= 1e-13 * np.abs(phis[ek.DC_ZERO]).max()


def test_startup_matches_the_oracle(ek):
    over = dict(NX=12, NY=4, NZ=11, pb_iters=60)
    o = eo.Oracle(eo.default_params(**over))
    o.set_poisson_dc(0)
    o.initialization()
    sim = ek.Simulation(ek.default_params(**over))
    sim.initialization()
    check(util.field_errors(sim.fields(), o.fields()))
    sim.init_equilibrium()
    o.init_equilibrium()
    for s in range(4):
        a, b = sim.populations(s), o.populations(s)
        assert np.abs(a - b).max() <= 1e-14 * np.abs(b).max()
    sim.close()


def test_product_library_has_no_cross_check_variants(ek):
    """libek_b200.so ships the hot path only: the slower kernel variants and the literal
    odd-extension transform live in libek_b200_xcheck.so (test infrastructure)"""
    assert ek.load_library().ek_is_xcheck_build() == 0
    assert ek.load_library(ek.XCHECK_LIB_PATH).ek_is_xcheck_build() == 1
    sim = ek.Simulation(ek.default_params(NX=8, NY=2, NZ=7))
    for key, value in (("kernel", 1), ("kernel", 2), ("kernel", 5), ("kernel", 6), ("poisson_path", 1)):
        with pytest.raises(ek.EkError):
            sim.set_option(key, value)
    with pytest.raises(ek.EkError):
        sim.set_poisson_dc(ek.DC_LITERAL)
    sim.set_option("kernel", 3)
    sim.close()


def test_step_graph_is_bitwise_identical(ek):
    """ek_step replays a CUDA graph of two coupled steps on small grids (option "graph"): same
    kernels, same arguments, so the same bits as the launch-by-launch loop"""
    over = dict(NX=40, NY=5, NZ=21, uw=1.0e-4, exf=1.0e6)
    init = synthetic_init(over)
    res = {}
    for mode in (ek.STREAM_AA, ek.STREAM_PUSH):
        for graph in (0, 1, -1):
            sim = ek.Simulation(ek.default_params(**over), stream_mode=mode)
            sim.set_option("graph", graph)
            sim.set_fields(init)
            sim.init_equilibrium()
            sim.step(11)
            sim.step(1)
            sim.step(8)
            replays = sim.counter("graph_replays")
            assert (replays == 0) if graph == 0 else (replays >= 6), (graph, replays)
            assert sim.counter("steps") == 20
            res[mode, graph] = (sim.fields(), np.stack([sim.populations(s) for s in range(4)]))
            sim.close()
    base_f, base_p = res[ek.STREAM_AA, 0]
    for key, (f, p) in res.items():
        for k in util.FIELDS:
            assert np.array_equal(f[k], base_f[k]), (key, k)
        assert np.array_equal(p, base_p), key


@pytest.mark.parametrize("nsteps", [1, 2, 3, 6])
def test_run_from_host_is_bitwise_identical_to_the_plain_sequence(ek, nsteps):
    """ek_run_from_host (upload, init_equilibrium, n steps, download as ONE call with the PCIe copies
    pipelined against the first and last LBM pass) against set_fields + init_equilibrium + step + fields"""
    over = dict(NX=40, NY=6, NZ=37, uw=1.0e-4, exf=1.0e6, voltage2=-3.0e-3)
    init = synthetic_init(over)
    sim = ek.Simulation(ek.default_params(**over), zchunk=4)
    sim.set_fields(init)
    sim.init_equilibrium()
    sim.step(nsteps)
    want = sim.fields()
    wantP = np.stack([sim.populations(s) for s in range(4)])
    sim.close()
    sim = ek.Simulation(ek.default_params(**over), zchunk=4)
    got = sim.run_from_host(init, nsteps)
    gotP = np.stack([sim.populations(s) for s in range(4)])
    assert sim.counter("steps") == nsteps
    for k in util.FIELDS:
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(gotP, wantP)
    # and again on the used handle, into caller-provided arrays
    out = {k: np.empty_like(want[k]) for k in util.FIELDS}
    sim.run_from_host(init, nsteps, out)
    for k in util.FIELDS:
        assert np.array_equal(out[k], want[k]), k
    sim.close()


def test_state_machine_errors(ek):
    sim = ek.Simulation(ek.default_params(NX=8, NY=2, NZ=7))
    with pytest.raises(ek.EkError):
        sim.step(1)                       # before any initialisation
    with pytest.raises(ek.EkError):
        sim.init_equilibrium()
    sim.set_fields({"rho": np.full(sim.shape, 1000.0)})
    sim.init_equilibrium()
    with pytest.raises(ek.EkError):
        sim.set_option("stream_mode", 1)  # storage already allocated
    with pytest.raises(ek.EkError):
        sim.set_option("zchunk", 1)       # the owner of z=0 must own z=1
    sim.step(2)
    assert sim.counter("steps") == 2
    sim.close()


# ---------------------------------------------------------------------------
# the reference's own CUDA build, run live (its binaries travel in oracle/_ref)
# ---------------------------------------------------------------------------
@pytest.mark.skipif(not util.have_ref("g2"), reason="oracle/_ref/ek_ref_g2 not built")
def test_live_reference_without_replay(ek):
    """NE = 32: the reference's forward cuFFT leaves an exactly zero DC
    coefficient on this grid, so no replay is needed; 100 coupled steps."""
    init, ref, _, info = util.run_ref("g2", 100, perturb=0.05, dc=True)
    assert np.all(info["dc"] == 0.0)
    got, _ = product_run(ek, util.case_overrides("g2"), init, 100, ek.STREAM_AA)
    err = util.field_errors(got, ref)
    _record("live_g2_100_steps_no_replay", err)
    check(err)


@pytest.mark.skipif(not util.have_ref("g3"), reason="oracle/_ref/ek_ref_g3 not built")
def test_live_reference_with_replay_moving_wall(ek):
    init, ref, refP, info = util.run_ref("g3", 120, perturb=0.05, pops=True, dc=True)
    got, P = product_run(ek, util.case_overrides("g3"), init, 120, ek.STREAM_AA, dc=info["dc"], pops=True)
    err = util.field_errors(got, ref)
    _record("live_g3_120_steps_replay", err)
    check(err)
    for s, e in enumerate(util.pop_errors(P, refP)):
        assert e <= 1e-12, (s, e)


# ---------------------------------------------------------------------------
# the BENCHMARKED shapes against the reference's CUDA build, with the 3-D perturbation of
# SURVEY.md 8(d) so that x/y streaming and every Fourier mode take part (main.cu:189-200)
# ---------------------------------------------------------------------------
@pytest.mark.skipif(not util.have_ref("c2"), reason="oracle/_ref/ek_ref_c2 not built")
def test_live_reference_c2_perturbed_200_steps(ek):
    """Config C2 (128x64x64, isothermal): the reference's own start-up, perturbed, 200 coupled
    steps; all 11 fields and the four population sets.  NE = 126 is not a power of two, so the
    reference's per-step (0,0,0) coefficient is replayed."""
    steps = 200
    init, ref, refP, info = util.run_ref("c2", steps, perturb=0.05, pops=True, dc=True)
    over = util.case_overrides("c2")
    got, P = product_run(ek, over, init, steps, ek.STREAM_AA, dc=info["dc"], pops=True)
    err = util.field_errors(got, ref)
    pe = util.pop_errors(P, refP)
    _record("live_c2_128x64x64_200_steps_replay", dict(err, **{f"pops_{s}": e for s, e in enumerate(pe)}))
    assert np.all(got["T"] == 0.0) and np.all(ref["T"] == 0.0)      # TH = 0: T stays exactly zero on both sides
    check(err)
    for s, e in enumerate(pe):
        assert e <= 1e-12, (s, e)
    # x/y structure really is there: the perturbation survives in c+ (not an x-y uniform comparison)
    assert np.abs(ref["charge"] - ref["charge"][:, :1, :1]).max() > 1e-3 * np.abs(ref["charge"]).max()


@pytest.mark.skipif(not util.have_ref("c3"), reason="oracle/_ref/ek_ref_c3 not built")
def test_live_reference_c3_perturbed_256cubed(ek):
    """Config C3 (256^3, the benchmarked grid): start-up by this library (0.4 s; the reference's
    takes 40 s), 3-D perturbation, then the SAME initial arrays go through the reference's CUDA
    build (--load-init) and through the product for 12 coupled steps with the DC replay; all 11
    fields at full size."""
    steps = 12
    over = util.case_overrides("c3")
    sim = ek.Simulation(ek.default_params(**over))
    sim.initialization()
    init = eo.perturb_fields(sim.fields(), 0.05)
    sim.close()
    _, ref, _, info = util.run_ref("c3", steps, init_fields=init, dc=True, dump_init=False)
    got, _ = product_run(ek, over, init, steps, ek.STREAM_AA, dc=info["dc"])
    err = util.field_errors(got, ref)
    _record("live_c3_256x256x256_12_steps_replay", err)
    check(err)
    assert np.abs(ref["charge"] - ref["charge"][:, :1, :1]).max() > 1e-3 * np.abs(ref["charge"]).max()
    # without the replay the difference is the reference's DC artefact and nothing else:
    # a constant shift of the interior potential
    got0, _ = product_run(ek, over, init, 1, ek.STREAM_AA)
    _, ref1, _, _ = util.run_ref("c3", 1, init_fields=init, dump_init=False)
    d = (got0["phi"] - ref1["phi"])[1:-1]
    _record("live_c3_dc_shift_1_step", {"shift_over_max_phi": float(np.abs(d).max() / np.abs(ref1["phi"]).max()),
                                        "spread_over_max_phi": float(np.ptp(d) / np.abs(ref1["phi"]).max())})
    assert np.ptp(d) <= 1e-12 * np.abs(ref1["phi"]).max()


# ---------------------------------------------------------------------------
# full-size, size-independent properties (config C3 256^3)
# ---------------------------------------------------------------------------
def test_full_size_properties(ek):
    # c_inf = 0.002: with the shipped 0.01 the reference's PB start-up diverges for NZ >~ 200
    over = dict(NX=256, NY=256, NZ=256, pb_iters=40, chargeinf=0.002)
    col = dict(NX=2, NY=2, NZ=256, pb_iters=40, chargeinf=0.002)
    steps = 6
    # (1) the un-perturbed problem is x-y uniform: the 256^3 run must equal the
    #     oracle's 2x2x256 column to round-off, and stay uniform
    o = eo.Oracle(eo.default_params(**col))
    o.set_poisson_dc(0)
    o.initialization()
    o.init_equilibrium()
    o.step(steps)
    want = o.fields()
    sim = ek.Simulation(ek.default_params(**over), stream_mode=ek.STREAM_AA)
    sim.init()
    m0 = None
    sim.step(steps)
    got = sim.fields()
    for k in ("rho", "charge", "chargen", "phi", "T", "Ez"):
        w = want[k][:, 0, 0]
        assert np.abs(got[k][:, 0, 0] - w).max() <= 1e-11 * np.abs(w).max(), k
        assert np.abs(got[k] - got[k][:, :1, :1]).max() <= 1e-11 * np.abs(w).max(), k
    w = want["ux"][:, 0, 0]
    assert np.abs(got["ux"][:, 0, 0] - w).max() <= 1e-12 * np.abs(w).max() + K_U * util.u_ulp(want)
    # (2) A-A and two-lattice push agree bit for bit at full size
    aa_rho, aa_phi, aa_ux = got["rho"], got["phi"], got["ux"]
    sim.close()
    sim = ek.Simulation(ek.default_params(**over), stream_mode=ek.STREAM_PUSH)
    sim.init()
    sim.step(steps)
    assert np.array_equal(sim.field("rho"), aa_rho)
    assert np.array_equal(sim.field("phi"), aa_phi)
    assert np.array_equal(sim.field("ux"), aa_ux)
    # (3) fluid mass is conserved (periodic + bounce-back)
    f = sim.populations(0)
    mass = f.sum(dtype=np.float64)
    assert abs(mass - 1000.0 * 256 ** 3) <= 1e-9 * 1000.0 * 256 ** 3
    sim.close()


# ---------------------------------------------------------------------------
# diagnostics and dumps
# ---------------------------------------------------------------------------
def test_diagnostics_and_dumps(ek, tmp_path):
    over = dict(NX=12, NY=4, NZ=9)
    init = synthetic_init(over)
    sim = ek.Simulation(ek.default_params(**over))
    sim.set_fields(init)
    sim.init_equilibrium()
    sim.step(5)
    f = sim.fields()
    p = sim.p
    # current(): LBM.cu:2674-2710
    c = 2.0 * f["charge"][-2] - f["charge"][-3]
    cn = 2.0 * f["chargen"][-2] - f["chargen"][-3]
    want = ((c - cn) * f["Ez"][-1]).sum() * p.K * p.dz * p.dz
    assert abs(sim.current() - want) <= 1e-12 * abs(want)
    # record_umax(): LBM.cu:2712-2753
    assert sim.max_uz() == max(0.0, float(f["uz"].max()))
    # save_data_tecplot / save_data_end: LBM.cu:2492-2627
    tec = tmp_path / "data.dat"
    sim.save_data_tecplot(str(tec), time=5e-10, append=False, first=True)
    lines = tec.read_text().splitlines()
    assert lines[0].startswith('VARIABLES="x","y","z","u","v","w","p","charge","neg charge","phi","Ex","Ey","Ez","Temperature"')
    assert lines[1] == ""
    assert lines[2] == 'ZONE T="t=5e-10", F=POINT, I = 12, J = 4, K = 9'
    assert len(lines) == 3 + 12 * 4 * 9
    ext = {k: f[k].copy() for k in f}
    for k in ("rho", "charge", "chargen", "ux", "uy", "uz"):
        ext[k][0] = 2.0 * f[k][1] - f[k][2]
        ext[k][-1] = 2.0 * f[k][-2] - f[k][-3]
    x, y, z = 5, 2, 0
    want_line = "%g %g %g %g %g %g %g %g %10.6f %10.6f %10.6f %10.6f %10.6f %10.6f" % (
        p.dx * x, p.dy * y, p.dz * z, ext["ux"][z, y, x], ext["uy"][z, y, x], ext["uz"][z, y, x], ext["rho"][z, y, x],
        ext["charge"][z, y, x], ext["chargen"][z, y, x], ext["phi"][z, y, x], ext["Ex"][z, y, x], ext["Ey"][z, y, x],
        ext["Ez"][z, y, x], ext["T"][z, y, x])
    assert lines[3 + (z * 4 + y) * 12 + x] == want_line
    end = tmp_path / "data_end.dat"
    sim.save_data_end(str(end), time=5e-10)
    el = end.read_text().splitlines()
    assert len(el) == 12 * 4 * 9
    assert len(el[0].split()) == 12
    sim.close()


# ---------------------------------------------------------------------------
# the two realisations of the Poisson stage (ek_poisson.cu)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("over", [dict(NX=16, NY=8, NZ=13), dict(NX=50, NY=6, NZ=33),
                                  dict(NX=8, NY=4, NZ=256, chargeinf=0.002),
                                  dict(NX=12, NY=5, NZ=64, voltage2=-1.0e-3)],
                         ids=lambda o: "x".join(str(o[k]) for k in ("NX", "NY", "NZ")))
def test_poisson_paths_agree(ek, over):
    """path 0 (2-D FFT + tridiagonal z-solve) against path 1 (the reference's
    odd-extension FFT) and against the oracle, on the same c+ - c-."""
    init = synthetic_init(over, pb_iters=40)
    o = eo.Oracle(eo.default_params(**over))
    o.set_poisson_dc(0)
    o.set_fields(init)
    o.fast_poisson()
    want = o.fields()
    got = {}
    for path in (0, 1):
        sim = ek.Simulation(ek.default_params(**over), xcheck=(path == 1))   # path 0 = the product library
        sim.set_option("poisson_path", path)
        sim.set_fields(init)
        sim.init_equilibrium()
        sim.fast_Poisson(True)
        got[path] = sim.fields()
        sim.close()
        for k in ("phi", "Ex", "Ey", "Ez"):
            scale = max(np.abs(want[n]).max() for n in (("phi",) if k == "phi" else ("Ex", "Ey", "Ez")))
            assert np.abs(got[path][k] - want[k]).max() <= 1e-12 * scale, (path, k)
        assert np.all(got[path]["phi"][0] == sim.p.voltage) and np.all(got[path]["phi"][-1] == sim.p.voltage2)
    assert np.abs(got[0]["phi"] - got[1]["phi"]).max() <= 1e-12 * np.abs(want["phi"]).max()


def test_prescribed_dc_is_the_same_shift_on_both_paths(ek):
    over = dict(NX=16, NY=4, NZ=12)
    init = synthetic_init(over)
    res = {}
    for path in (0, 1):
        sim = ek.Simulation(ek.default_params(**over), xcheck=(path == 1))
        sim.set_option("poisson_path", path)
        sim.set_fields(init)
        sim.init_equilibrium()
        sim.set_poisson_dc(ek.DC_PRESCRIBED, 0.37)
        sim.fast_Poisson(True)
        res[path] = sim.field("phi")
        sim.set_poisson_dc(ek.DC_ZERO)
        sim.fast_Poisson(True)
        base = sim.field("phi")
        size = 16 * 4 * 2 * 11
        assert np.abs((res[path] - base)[1:-1] + 0.37 / size).max() < 1e-16
        sim.close()
    assert np.abs(res[0] - res[1]).max() <= 1e-13 * np.abs(res[1]).max()


# ---------------------------------------------------------------------------
# the reference-signature shim (libek_b200_shim.so)
# ---------------------------------------------------------------------------
def test_reference_signature_shim(ek):
    """Drive the library exactly as the reference's main() drives its own code:
    caller-owned device arrays, init_equilibrium(18 ptrs), then
    stream_collide_save(24 args) + fast_Poisson(6 args) per step."""
    import ctypes as C
    import subprocess
    import torch
    so = os.path.join(os.path.dirname(ek.LIB_PATH), "libek_b200_shim.so")
    syms = subprocess.run(["nm", "-D", so], capture_output=True, text=True, check=True).stdout.split()
    mangled = {n: next(s for s in syms if s.startswith("_Z") and n in s) for n in
               ("init_equilibrium", "stream_collide_save", "fast_Poisson")}
    L = C.CDLL(so)
    over = dict(NX=32, NY=4, NZ=13)
    init = synthetic_init(over)
    p = ek.default_params(**over)
    L.ek_shim_configure(C.byref(p))
    dev = torch.device("cuda", 0)
    arr = {k: torch.from_numpy(np.ascontiguousarray(init[k])).to(dev) for k in util.FIELDS}
    ptr = {k: C.c_void_p(v.data_ptr()) for k, v in arr.items()}
    scratch = torch.zeros(8, dtype=torch.float64, device=dev)  # stands in for the caller's population arrays
    sp = C.c_void_p(scratch.data_ptr())
    L.ek_shim_bind_potential(ptr["phi"], ptr["Ex"], ptr["Ey"], ptr["Ez"])
    getattr(L, mangled["init_equilibrium"])(sp, sp, sp, sp, sp, sp, sp, sp, ptr["rho"], ptr["charge"], ptr["chargen"],
                                            ptr["ux"], ptr["uy"], ptr["uz"], ptr["Ex"], ptr["Ey"], ptr["Ez"], ptr["T"])
    scs = getattr(L, mangled["stream_collide_save"])
    scs.argtypes = [C.c_void_p] * 22 + [C.c_double, C.c_void_p]
    fp = getattr(L, mangled["fast_Poisson"])
    fp.argtypes = [C.c_void_p] * 5 + [C.c_int]
    steps = 5
    for _ in range(steps):
        scs(sp, sp, sp, sp, sp, sp, sp, sp, sp, sp, sp, sp, ptr["rho"], ptr["charge"], ptr["chargen"], ptr["ux"],
            ptr["uy"], ptr["uz"], ptr["Ex"], ptr["Ey"], ptr["Ez"], ptr["T"], 0.0, sp)
        fp(ptr["charge"], ptr["chargen"], sp, sp, sp, 0)
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in arr.items()}
    want, _ = oracle_run(over, init, steps)
    check(util.field_errors(got, want))


@pytest.mark.parametrize("NX", [40, 64, 96, 128, 256])
def test_kernel_variants_agree(ek, NX):
    """LBM kernel variants: 0 (default: z-walking CTAs, lean deep-interior path; for NX = 128, 256 (4, 8 x-tiles)
    the odd step takes the instantiation with the row stride as an immediate), 4 (the generic lean kernel for
    every row length), 5 / 6 (x-marching rows
    for the odd A-A step when NX % 32 == 0: sector-aligned stores / aligned loads too), 3 (general node path
    everywhere) and the cross-check build's five-warp kernel 2 chain the sums in the reference's
    order and must agree bit for bit; the eight-warp kernel 1 adds two partial sums."""
    over = dict(NX=NX, NY=5, NZ=21, uw=1.0e-4, exf=1.0e6)
    init = synthetic_init(over)
    res = {}
    for kernel in (0, 1, 2, 3, 4, 5, 6):
        for mode in (ek.STREAM_AA, ek.STREAM_PUSH):
            sim = ek.Simulation(ek.default_params(**over), stream_mode=mode, zchunk=6, xcheck=kernel in (1, 2, 5, 6))
            sim.set_option("kernel", kernel)
            sim.set_option("graph", 0)
            sim.set_fields(init)
            sim.init_equilibrium()
            sim.step(7)
            res[kernel, mode] = (sim.fields(), np.stack([sim.populations(s) for s in range(4)]))
            sim.close()
    base_f, base_p = res[3, ek.STREAM_AA]
    for key in res:
        if key[0] == 1:
            continue
        f, p = res[key]
        for k in util.FIELDS:
            assert np.array_equal(f[k], base_f[k]), (key, k)
        assert np.array_equal(p, base_p), key
    for key in ((1, ek.STREAM_AA), (1, ek.STREAM_PUSH)):
        check(util.field_errors(res[key][0], base_f))
    want, _ = oracle_run(over, init, 7)
    for key in res:
        check(util.field_errors(res[key][0], want))


def test_marching_odd_step_in_z_ranges_and_with_fields(ek):
    """the x-marching odd step launched per z-block range (as the slab pipeline does) and with the
    macroscopic arrays written on odd steps: bit-identical to the z-walking kernel"""
    over = dict(NX=64, NY=6, NZ=23, exf=1.0e6)
    init = synthetic_init(over)
    res = []
    for kernel in (0, 5, 6):
        sim = ek.Simulation(ek.default_params(**over), zchunk=4, xcheck=kernel >= 5)
        sim.set_option("kernel", kernel)
        sim.set_fields(init)
        sim.init_equilibrium()
        nb = -(-23 // 4)
        for step in range(6):
            cuts = [0, 1, 3, nb] if kernel >= 5 else [0, nb]
            for b0, b1 in zip(cuts[:-1], cuts[1:]):
                sim._ck(sim.L.ek_stream_collide_save_range(sim.h, 1, b0, b1, int(b1 == nb)), "range")
            sim.fast_Poisson(True)
        res.append((sim.fields(), np.stack([sim.populations(s) for s in range(4)])))
        sim.close()
    for other in res[1:]:
        for k in util.FIELDS:
            assert np.array_equal(res[0][0][k], other[0][k]), k
        assert np.array_equal(res[0][1], other[1])


# ---------------------------------------------------------------------------
# config C2: isothermal electro-osmotic slit flow against the analytic profiles
# ---------------------------------------------------------------------------
def test_c2_slit_flow_matches_the_analytic_profiles(ek):
    """BASELINE config C2 (128x64x64, Ext = 1e4 V/m, TH = 0, no body force) run to the
    steady state (8000 steps = 7 viscous times (H/2)^2/nu of the half channel):
      * potential: Debye-Hueckel profile between two walls at zeta,
            phi = zeta cosh(kappa (z - H/2)) / cosh(kappa H/2), kappa^2 = 2 F c_inf e / (eps kB T0)
        (zeta = 5.3 mV << kB T0/e = 23.5 mV, so the linearisation error is ~3e-3);
      * velocity: Helmholtz-Smoluchowski, u_x = eps Ext (phi - phi_slip) / (rho0 nu).  The reference's
        full-way bounce-back (LBM.cu:1862-1887) puts the no-slip plane half a cell inside the wall
        node while the Dirichlet value of phi sits on the node (poisson.cu:195-201), hence phi_slip
        = phi(dz/2).  Measured: 2.4e-3 (phi), 2.2e-3 (u vs simulated phi), 3.6e-3 (u vs analytic phi)."""
    p = ek.default_params(NX=128, NY=64, NZ=64, TH=0.0, exf=0.0, Ext=1.0e4)
    sim = ek.Simulation(p)
    sim.init()
    sim.step(8000)
    phi3, ux3, T3 = sim.field("phi"), sim.field("ux"), sim.field("T")
    sim.close()
    NZ = p.NZ
    z = np.arange(NZ) * p.dz
    H = (NZ - 1) * p.dz
    zeta = p.voltage
    kappa = np.sqrt(2.0 * p.convertCtoCharge * p.chargeinf * p.electron / (p.eps * p.kB * p.roomT))
    phi_dh = zeta * np.cosh(kappa * (z - 0.5 * H)) / np.cosh(0.5 * kappa * H)
    phi, ux = phi3[:, 0, 0], ux3[:, 0, 0]
    mob = p.eps * p.Ext / (p.rho0 * p.nu)
    u_scale = abs(mob * zeta)
    assert np.abs(T3).max() == 0.0                                        # isothermal: T stays exactly 0
    assert np.abs(ux3 - ux3[:, :1, :1]).max() <= 1e-9 * u_scale           # x-y uniform
    assert np.abs(phi - phi_dh).max() <= 1e-2 * abs(zeta)
    inner = slice(1, NZ - 1)
    phi_slip = 0.25 * (phi[0] + phi[1] + phi[-1] + phi[-2])
    assert np.abs(ux[inner] - mob * (phi[inner] - phi_slip)).max() <= 1e-2 * u_scale
    phi_slip_dh = zeta * np.cosh(kappa * (0.5 * p.dz - 0.5 * H)) / np.cosh(0.5 * kappa * H)
    assert np.abs(ux[inner] - mob * (phi_dh[inner] - phi_slip_dh)).max() <= 1e-2 * u_scale


# ---------------------------------------------------------------------------
# restart: read_data() text format and the exact checkpoint
# ---------------------------------------------------------------------------
def test_checkpoint_continues_bit_for_bit(ek, tmp_path):
    """save after 5 steps (A-A parity 1), load into a fresh handle (natural layout, either
    streaming scheme), continue 4 steps: identical to the uninterrupted 9-step run"""
    over = dict(NX=40, NY=5, NZ=13, uw=1.0e-4, exf=1.0e6)
    init = synthetic_init(over)
    sim = ek.Simulation(ek.default_params(**over))
    sim.set_fields(init)
    sim.init_equilibrium()
    sim.step(5)
    path = str(tmp_path / "state.ekc")
    sim.checkpoint_save(path)
    sim.step(4)
    want_f, want_p = sim.fields(), np.stack([sim.populations(s) for s in range(4)])
    sim.close()
    for mode in (ek.STREAM_AA, ek.STREAM_PUSH):
        sim = ek.Simulation(ek.default_params(**over), stream_mode=mode)
        t = sim.checkpoint_load(path)
        assert abs(t - 5 * sim.p.dt) <= 1e-20
        sim.step(4)
        got_f, got_p = sim.fields(), np.stack([sim.populations(s) for s in range(4)])
        sim.close()
        for k in util.FIELDS:
            assert np.array_equal(got_f[k], want_f[k]), (mode, k)
        assert np.array_equal(got_p, want_p), mode
    # a checkpoint of another grid is refused
    sim = ek.Simulation(ek.default_params(NX=40, NY=5, NZ=15))
    with pytest.raises(ek.EkError):
        sim.checkpoint_load(path)
    sim.close()


def test_read_data_restores_the_reference_text_restart(ek, tmp_path):
    """save_data_end() -> read_data() (LBM.cu:2567-2671): the file keeps six decimals and the
    dump-time wall extrapolation, so the restored arrays equal the dumped ones to 5e-7 absolute"""
    over = dict(NX=12, NY=4, NZ=9)
    sim = ek.Simulation(ek.default_params(**dict(over, pb_iters=30)))
    sim.init()
    sim.step(3)
    path = str(tmp_path / "data_end.dat")
    sim.save_data_end(path, 3.0e-10)
    dumped = sim.fields()
    sim.close()
    sim = ek.Simulation(ek.default_params(**over))
    t = sim.read_data(path)
    assert abs(t - 0.0) <= 5e-7          # "%10.6f" of 3e-10
    got = sim.fields()
    sim.init_equilibrium()               # main.cu:174 -- the restart re-creates the populations
    sim.step(1)
    sim.close()
    for k in ("phi", "T", "Ex", "Ey", "Ez"):
        assert np.abs(got[k] - dumped[k]).max() <= 5.1e-7, k
    for k in ("rho", "charge", "chargen", "ux", "uy", "uz"):
        a = dumped[k].copy()                                  # LBM.cu:2598-2613
        a[0] = 2.0 * a[1] - a[2]
        a[-1] = 2.0 * a[-2] - a[-3]
        assert np.abs(got[k] - a).max() <= 5.1e-7, k
    with pytest.raises(ek.EkError):
        ek.Simulation(ek.default_params(NX=12, NY=4, NZ=11)).read_data(path)   # too few cells


# ---------------------------------------------------------------------------
# the reference's main() over the C ABI (examples/ek_main.cpp)
# ---------------------------------------------------------------------------
def test_cpp_main_matches_the_python_host(ek, tmp_path):
    """examples/ek_main.cpp = main.cu:19-295 on the C ABI: same flow (stdin prompt, start-up,
    loop with dumps and diagnostics, data_end.dat), compared with the same run driven from Python"""
    import subprocess
    exe = os.path.join(util.ROOT, "ek-pnp-3d_b200", "ek_main")
    if not os.path.exists(exe):
        pytest.skip("ek_main not built")
    args = ["--nx", "12", "--ny", "4", "--nz", "9", "--nsteps", "24", "--nsave", "10", "--print-current", "5",
            "--pb-iters", "30", "--checkpoint", "state.ekc"]
    out = subprocess.run([exe] + args, input="0\n", capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "Initializing..." in out.stdout and "speed:" in out.stdout and "Current =" in out.stdout
    for f in ("data.dat", "umax.dat", "data_end.dat", "state.ekc"):
        assert (tmp_path / f).stat().st_size > 0, f
    assert open(tmp_path / "data.dat").read().count("ZONE T=") == 5          # t = 0, i = 1, 11, 21 (i % NSAVE == 1), end
    over = dict(NX=12, NY=4, NZ=9, pb_iters=30)
    sim = ek.Simulation(ek.default_params(**over))
    sim.init()
    sim.step(24)
    want = sim.fields()
    sim.close()
    raw = np.loadtxt(tmp_path / "data_end.dat").reshape(9, 4, 12, 12)
    cols = {"ux": 1, "uy": 2, "uz": 3, "rho": 4, "charge": 5, "chargen": 6, "phi": 7, "Ex": 8, "Ey": 9, "Ez": 10, "T": 11}
    for k, j in cols.items():
        a = want[k].copy()
        if k in ("rho", "charge", "chargen", "ux", "uy", "uz"):     # dump-time wall extrapolation, LBM.cu:2598-2613
            a[0] = 2.0 * a[1] - a[2]
            a[-1] = 2.0 * a[-2] - a[-3]
        assert np.abs(raw[..., j] - a).max() <= 5.1e-7, k
    # the checkpoint written by the C++ host resumes in the Python host
    sim = ek.Simulation(ek.default_params(**over))
    sim.checkpoint_load(str(tmp_path / "state.ekc"))
    for k in ("rho", "phi", "T"):
        assert np.array_equal(sim.field(k), want[k]), k
    sim.close()
    # the same run split into two x-slabs by the native multi-GPU driver (devices wrap on a 1-GPU box)
    (tmp_path / "two").mkdir()
    out2 = subprocess.run([exe] + args[:-2] + ["--gpus", "2"], capture_output=True, text=True, cwd=tmp_path / "two", timeout=120)
    assert out2.returncode == 0, out2.stderr
    raw2 = np.loadtxt(tmp_path / "two" / "data_end.dat").reshape(9, 4, 12, 12)
    assert np.abs(raw2 - raw).max() <= 1.1e-6        # one unit of the six-decimal text format
    assert open(tmp_path / "two" / "umax.dat").read().count("\n") == open(tmp_path / "umax.dat").read().count("\n")
    # restart from the text file, as the reference's "press 1" branch
    out = subprocess.run([exe, "--nx", "12", "--ny", "4", "--nz", "9", "--nsteps", "2"], input="1\n", capture_output=True,
                         text=True, cwd=tmp_path, timeout=120)
    assert out.returncode == 0 and "Reading previous data..." in out.stdout, out.stderr
