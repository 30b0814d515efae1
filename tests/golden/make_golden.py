#!/usr/bin/env python3
"""Generate the golden fixtures from the reference's own CUDA build.

Run on a GPU box (the reference is CUDA-only):
    gpurun -- python tests/golden/make_golden.py
It executes oracle/_ref/ek_ref_<case> (built by oracle/build_ref.py from the
unmodified sources under /root/reference) and writes gpurun_out/golden/*.npz,
which are then committed under tests/golden/.

Each fixture holds the raw fp64 macroscopic arrays of the reference before the
first step ("init_*": after initialization() [+ perturbation], i.e. the input
of init_equilibrium()) and after `steps` loop iterations ("final_*"), the
per-step (0,0,0) coefficient of the reference's forward cuFFT ("dc", see
DESIGN.md "DC artefact"), and -- for the un-perturbed start-up fixture -- the
state right after the reference's initialization().
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402

# case, perturbation amplitude, steps
FIXTURES = [("g1", 0.05, 50), ("g3", 0.05, 50), ("g4", 0.05, 50)]


def main():
    out = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out, exist_ok=True)
    for case, amp, steps in FIXTURES:
        init, final, pops, info = util.run_ref(case, steps, perturb=amp, pops=True, dc=True)
        init2, final2, _, info2 = util.run_ref(case, steps, perturb=amp, dc=True)
        assert all(np.array_equal(final[k], final2[k]) for k in final), "reference not run-to-run stable"
        arrays = {f"init_{k}": v for k, v in init.items()}
        arrays.update({f"final_{k}": v for k, v in final.items()})
        arrays["dc"] = info["dc"]
        # velocity-sensitive check: the fluid populations themselves (z-profile at x=y=0 is not enough)
        arrays["final_fluid_pops"] = pops[0]
        meta = {"case": case, "perturb": amp, "steps": steps, "overrides": util.case_overrides(case),
                "reference_info": {k: v for k, v in info.items() if k != "dc"}}
        arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        np.savez_compressed(os.path.join(out, f"{case}_p{int(amp * 100):02d}_s{steps}.npz"), **arrays)
        print("wrote", case, steps, "dc range", float(info["dc"].min()), float(info["dc"].max()))
    # start-up fixture: the reference's initialization() on a grid where its DC coefficient is exactly zero
    init, _, _, _ = util.run_ref("g4", 0)
    meta = {"case": "g4", "what": "fields after the reference's initialization() (LBM.cu:68-146)",
            "overrides": util.case_overrides("g4")}
    arrays = {f"init_{k}": v for k, v in init.items()}
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(out, "g4_startup.npz"), **arrays)
    # shipped case C1 (50x8x51, 1000 steps, no perturbation): x-y uniform, keep the z profile at x=y=0
    init, final, _, info = util.run_ref("c1", 1000, dc=True)
    arrays = {}
    for k in final:
        arrays[f"final_{k}_zprofile"] = final[k][:, 0, 0].copy()
        arrays[f"final_{k}_xy_spread"] = np.array(float(np.abs(final[k] - final[k][:, :1, :1]).max()))
        arrays[f"init_{k}_zprofile"] = init[k][:, 0, 0].copy()
    arrays["dc"] = info["dc"]
    meta = {"case": "c1", "steps": 1000, "overrides": util.case_overrides("c1"),
            "what": "shipped case; z profiles at x=y=0 plus the max deviation over x,y"}
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(out, "c1_shipped_s1000.npz"), **arrays)
    print("wrote c1")


if __name__ == "__main__":
    main()
