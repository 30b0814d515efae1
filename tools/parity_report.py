#!/usr/bin/env python3
"""Three-way parity report on a GPU box: reference CUDA build (oracle/_ref),
CPU restatement (oracle/), and the B200-native library.  Prints one JSON line
per comparison; nothing is asserted here (tests/ hold the tolerances).

"replay" runs feed the per-step (0,0,0) coefficient recorded from the
reference's own forward cuFFT into the Poisson stage of the oracle and of the
product ("DC artefact"), which removes the one
implementation-dependent number of the reference from the comparison."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import util  # noqa: E402
from oracle import ek_oracle as eo  # noqa: E402


def run_product(ek, case, init, steps, mode, dc=None, dc_mode=None, zchunk=None, pops=False):
    sim = ek.Simulation(util.product_params(case), stream_mode=mode, zchunk=zchunk)
    sim.set_fields(init)
    sim.init_equilibrium()
    if dc is not None:
        for n in range(steps):
            sim.set_poisson_dc(ek.DC_PRESCRIBED, float(dc[n]))
            sim.step(1)
    else:
        if dc_mode is not None:
            sim.set_poisson_dc(dc_mode)
        sim.step(steps)
    f = sim.fields()
    P = np.stack([sim.populations(s) for s in range(4)]) if pops else None
    sim.close()
    return f, P


def run_oracle(case, init, steps, dc=None, dc_mode=None, pops=False):
    o = eo.Oracle(util.oracle_params(case))
    o.set_fields(init)
    o.init_equilibrium()
    if dc is not None:
        for n in range(steps):
            o.set_poisson_dc(2, float(dc[n]))
            o.step(1)
    else:
        if dc_mode is not None:
            o.set_poisson_dc(dc_mode)
        o.step(steps)
    f = o.fields()
    P = np.stack([o.populations(s) for s in range(4)]) if pops else None
    o.close()
    return f, P


def main():
    ek = util.ek_module()
    out = []
    cases = [("g1", 0.05, (1, 2, 10, 100, 1000)), ("g2", 0.05, (1, 10, 100)), ("g3", 0.05, (1, 2, 10, 100)),
             ("c1", 0.0, (10, 1000))]
    if len(sys.argv) > 1:
        cases = [c for c in cases if c[0] in sys.argv[1:]]
    for case, amp, step_list in cases:
        if not util.have_ref(case):
            print(json.dumps({"case": case, "skip": "no reference binary"}))
            continue
        for steps in step_list:
            t0 = time.time()
            init, ref, refP, info = util.run_ref(case, steps, perturb=amp, pops=True, dc=True)
            dc = info["dc"]
            rec = {"case": case, "steps": steps, "perturb": amp,
                   "ref_dc_min_max": [float(dc.min()), float(dc.max())]}
            # literal oracle (its own FFT residue) vs reference: bounded by the DC artefact
            orc, _ = run_oracle(case, init, steps)
            rec["oracle_literal_vs_ref"] = util.field_errors(orc, ref)
            # replayed DC: everything else must agree tightly
            orc_r, orcP_r = run_oracle(case, init, steps, dc=dc, pops=True)
            rec["oracle_replay_vs_ref"] = util.field_errors(orc_r, ref)
            rec["oracle_replay_vs_ref_pops"] = util.pop_errors(orcP_r, refP)
            orc_z, _ = run_oracle(case, init, steps, dc_mode=0)
            for mode, name in ((ek.STREAM_AA, "aa"), (ek.STREAM_PUSH, "push")):
                got, gotP = run_product(ek, case, init, steps, mode, dc=dc, pops=True)
                rec[f"{name}_replay_vs_ref"] = util.field_errors(got, ref)
                rec[f"{name}_replay_vs_ref_pops"] = util.pop_errors(gotP, refP)
                rec[f"{name}_replay_vs_oracle_replay"] = util.field_errors(got, orc_r)
                gz, _ = run_product(ek, case, init, steps, mode, dc_mode=ek.DC_ZERO)
                rec[f"{name}_zero_vs_oracle_zero"] = util.field_errors(gz, orc_z)
                rec[f"{name}_zero_vs_ref"] = util.field_errors(gz, ref)
            if case == "g2" and util.have_ref("g2_nofmad"):
                _, ref2, ref2P, _ = util.run_ref("g2_nofmad", steps, init_fields=init, pops=True)
                rec["ref_nofmad_vs_ref"] = util.field_errors(ref2, ref)
                rec["ref_nofmad_vs_ref_pops"] = util.pop_errors(ref2P, refP)
            # run-to-run stability of the reference (race of SURVEY.md A.7-2)
            _, ref_again, _, _ = util.run_ref(case, steps, perturb=amp)
            rec["ref_rerun_bitwise_equal"] = all(np.array_equal(ref[k], ref_again[k]) for k in ref)
            rec["seconds"] = round(time.time() - t0, 2)
            print(json.dumps(rec), flush=True)
            out.append(rec)
        # start-up parity: initialization() on all three sides (unperturbed)
        init_ref, _, _, _ = util.run_ref(case, 0)
        rec = {"case": case, "init": True}
        for dcm, nm in ((1, "literal"), (0, "zero")):
            o = eo.Oracle(util.oracle_params(case)); o.set_poisson_dc(dcm); o.initialization(); io = o.fields(); o.close()
            sim = ek.Simulation(util.product_params(case)); sim.set_poisson_dc(dcm); sim.initialization()
            ip = sim.fields(); sim.close()
            rec[f"oracle_{nm}_vs_ref"] = util.field_errors(io, init_ref)
            rec[f"product_{nm}_vs_ref"] = util.field_errors(ip, init_ref)
            rec[f"product_{nm}_vs_oracle_{nm}"] = util.field_errors(ip, io)
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
