#!/usr/bin/env python3
"""LBM pass of ONE narrow x-slab (ghost columns, slab layout of c+ - c-) against the same grid as a plain
single domain: where does the slab mode lose at 128 columns?  (development aid, one GPU)"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401


def main():
    ek = importlib.import_module("ek-pnp-3d_b200")
    NX = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    over = dict(NX=NX, NY=256, NZ=256, pb_iters=20, chargeinf=0.002, exf=2.0e6)
    sim = ek.Simulation(ek.default_params(**over), profile=True)
    sim.init()
    sim.step(6)
    sim.reset_counters()
    sim.step(20)
    sim.sync()
    plain = {"lbm_ms": sim.counter("lbm_ms") / 20, "even": sim.counter("lbm_ms_even") / 10, "odd": sim.counter("lbm_ms_odd") / 10}
    sim.close()
    rs = ek.RankSimulation(ek.default_params(**over), 0, 0, 1, None, poisson_chunks=1)
    rs.init()
    rs.step(6)
    seq = rs.profile(20, True)
    rs.close()
    print(json.dumps({"NX": NX, "plain_single_domain": plain, "one_rank_slab_sequential": seq}), flush=True)


if __name__ == "__main__":
    main()
