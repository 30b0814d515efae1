#!/usr/bin/env python3
"""C1 / C2 coupled-step time with and without the step graph (development aid; bench.py reports
the same figures in its `extra` record)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402


def main():
    ek = util.ek_module()
    for name, over, steps in (("c1", dict(NX=50, NY=8, NZ=51), 1000), ("c2", dict(NX=128, NY=64, NZ=64, TH=0.0), 400)):
        for graph in (0, 1):
            sim = ek.Simulation(ek.default_params(**over))
            sim.set_option("graph", graph)
            sim.init()
            sim.step(20)
            best = min(sim.step_timed(steps) for _ in range(3))
            cells = over["NX"] * over["NY"] * over["NZ"]
            print(json.dumps({"case": name, "graph": graph, "us_per_step": round(1e3 * best / steps, 2),
                              "mlups": round(cells * steps / best / 1e3, 1), "zchunk": sim.counter("zchunk"),
                              "replays": sim.counter("graph_replays")}), flush=True)
            sim.close()


if __name__ == "__main__":
    main()
