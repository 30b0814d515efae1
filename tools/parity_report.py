#!/usr/bin/env python3
"""Three-way parity report on a GPU box: reference CUDA build (oracle/_ref),
CPU restatement (oracle/), and the B200-native library.  Prints one JSON line
per comparison; nothing is asserted here (tests/ hold the tolerances)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tests import util  # noqa: E402
from oracle import ek_oracle as eo  # noqa: E402


def run_product(ek, case, init, steps, mode, zchunk=None, pops=False):
    sim = ek.Simulation(util.product_params(case), stream_mode=mode, zchunk=zchunk)
    sim.set_fields(init)
    sim.init_equilibrium()
    sim.step(steps)
    f = sim.fields()
    P = np.stack([sim.populations(s) for s in range(4)]) if pops else None
    sim.close()
    return f, P


def run_oracle(case, init, steps, pops=False):
    o = eo.Oracle(util.oracle_params(case))
    o.set_fields(init)
    o.init_equilibrium()
    o.step(steps)
    f = o.fields()
    P = np.stack([o.populations(s) for s in range(4)]) if pops else None
    o.close()
    return f, P


def main():
    ek = util.ek_module()
    out = []
    cases = [("g1", 0.05, (1, 2, 3, 10, 100, 1000)), ("g2", 0.05, (1, 10, 100)), ("g3", 0.05, (1, 2, 10, 100)),
             ("c1", 0.0, (10, 1000))]
    if len(sys.argv) > 1:
        cases = [c for c in cases if c[0] in sys.argv[1:]]
    for case, amp, step_list in cases:
        if not util.have_ref(case):
            print(json.dumps({"case": case, "skip": "no reference binary"}))
            continue
        for steps in step_list:
            t0 = time.time()
            init, ref, refP, info = util.run_ref(case, steps, perturb=amp, pops=True)
            rec = {"case": case, "steps": steps, "perturb": amp}
            orc, orcP = run_oracle(case, init, steps, pops=True)
            rec["oracle_vs_ref"] = util.field_errors(orc, ref)
            rec["oracle_vs_ref_pops"] = util.pop_errors(orcP, refP)
            for mode, name in ((ek.STREAM_AA, "aa"), (ek.STREAM_PUSH, "push")):
                got, gotP = run_product(ek, case, init, steps, mode, pops=True)
                rec[f"{name}_vs_ref"] = util.field_errors(got, ref)
                rec[f"{name}_vs_oracle"] = util.field_errors(got, orc)
                rec[f"{name}_vs_ref_pops"] = util.pop_errors(gotP, refP)
            if case == "g2" and util.have_ref("g2_nofmad"):
                _, ref2, ref2P, _ = util.run_ref("g2_nofmad", steps, init_fields=init, pops=True)
                rec["ref_nofmad_vs_ref"] = util.field_errors(ref2, ref)
                rec["ref_nofmad_vs_ref_pops"] = util.pop_errors(ref2P, refP)
                rec["oracle_vs_ref_nofmad"] = util.field_errors(orc, ref2)
            # run-to-run stability of the reference (race of SURVEY.md A.7-2)
            _, ref_again, _, _ = util.run_ref(case, steps, perturb=amp)
            rec["ref_rerun_bitwise_equal"] = all(np.array_equal(ref[k], ref_again[k]) for k in ref)
            rec["seconds"] = round(time.time() - t0, 2)
            print(json.dumps(rec), flush=True)
            out.append(rec)
        # start-up parity: initialization() on all three sides (unperturbed)
        init_ref, _, _, _ = util.run_ref(case, 0)
        o = eo.Oracle(util.oracle_params(case)); o.initialization(); io = o.fields(); o.close()
        sim = ek.Simulation(util.product_params(case)); sim.initialization(); ip = sim.fields(); sim.close()
        rec = {"case": case, "init": True, "oracle_vs_ref": util.field_errors(io, init_ref),
               "product_vs_ref": util.field_errors(ip, init_ref)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
