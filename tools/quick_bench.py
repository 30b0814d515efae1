#!/usr/bin/env python3
"""Quick single-GPU timing sweeps of the coupled step (development aid, not bench.py).

usage: quick_bench.py NX NY NZ steps key=v1,v2 key=v1,v2 ...
keys: mode (aa|push), zchunk, prefetch, poisson (option values of ek_set_option)
"""
import itertools
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402


def main():
    ek = util.ek_module()
    NX, NY, NZ = (int(v) for v in sys.argv[1:4])
    steps = int(sys.argv[4])
    sweeps = {}
    for arg in sys.argv[5:]:
        k, v = arg.split("=")
        sweeps[k] = v.split(",")
    keys = list(sweeps)
    for combo in itertools.product(*[sweeps[k] for k in keys]):
        cfg = dict(zip(keys, combo))
        mode = ek.STREAM_PUSH if cfg.get("mode", "aa") == "push" else ek.STREAM_AA
        p = ek.default_params(NX=NX, NY=NY, NZ=NZ, pb_iters=20, chargeinf=0.002 if NZ > 160 else 0.01)
        sim = ek.Simulation(p, stream_mode=mode, profile=True)
        for k, v in cfg.items():
            if k != "mode":
                sim.set_option(k, int(v))
        sim.init()
        sim.step(6)
        sim.sync()
        sim.reset_counters()
        t0 = time.time()
        sim.step(steps)
        sim.sync()
        wall = time.time() - t0
        lbm = sim.counter("lbm_ms") / steps
        poi = sim.counter("poisson_ms") / steps
        even = sim.counter("lbm_ms_even") / max(1, (steps + 1) // 2)
        odd = sim.counter("lbm_ms_odd") / max(1, steps // 2)
        cells = NX * NY * NZ
        print(json.dumps({**cfg, "grid": [NX, NY, NZ], "ms_per_step_wall": round(1e3 * wall / steps, 4),
                          "lbm_ms": round(lbm, 4), "lbm_even_ms": round(even, 4), "lbm_odd_ms": round(odd, 4), "poisson_ms": round(poi, 4),
                          "mlups_wall": round(cells * steps / wall / 1e6, 1),
                          "lbm_GBps_alg": round(cells * 1744 / (lbm * 1e-3) / 1e9, 1)}), flush=True)
        sim.close()


if __name__ == "__main__":
    main()
