#!/usr/bin/env python3
"""Quick single-GPU timing of the coupled step (development aid, not bench.py)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402


def main():
    ek = util.ek_module()
    NX, NY, NZ = (int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (256, 256, 256)))
    steps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    for mode, name in ((ek.STREAM_AA, "aa"), (ek.STREAM_PUSH, "push")):
        for zchunk in (4, 8, 16, 32):
            p = ek.default_params(NX=NX, NY=NY, NZ=NZ, pb_iters=20)
            sim = ek.Simulation(p, stream_mode=mode, zchunk=zchunk, profile=True)
            t0 = time.time()
            sim.init()
            sim.sync()
            t_init = time.time() - t0
            sim.step(4)
            sim.sync()
            sim.reset_counters()
            t0 = time.time()
            sim.step(steps)
            sim.sync()
            wall = time.time() - t0
            lbm = sim.counter("lbm_ms") / steps
            poi = sim.counter("poisson_ms") / steps
            cells = NX * NY * NZ
            print(json.dumps({"mode": name, "zchunk": zchunk, "grid": [NX, NY, NZ], "init_s": round(t_init, 3),
                              "ms_per_step_wall": round(1e3 * wall / steps, 4), "lbm_ms": round(lbm, 4),
                              "poisson_ms": round(poi, 4), "mlups_wall": round(cells * steps / wall / 1e6, 1),
                              "lbm_GBps_alg": round(cells * 1728 / (lbm * 1e-3) / 1e9, 1)}), flush=True)
            sim.close()


if __name__ == "__main__":
    main()
