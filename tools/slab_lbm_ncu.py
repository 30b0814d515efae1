#!/usr/bin/env python3
"""a few LBM launches of one narrow slab (argv[1] = slab) or of the same grid as a plain domain (plain), for ncu"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: F401
ek = importlib.import_module("ek-pnp-3d_b200")
over = dict(NX=128, NY=256, NZ=256, pb_iters=5, chargeinf=0.002, exf=2.0e6)
if sys.argv[1] == "slab":
    rs = ek.RankSimulation(ek.default_params(**over), 0, 0, 1, None, poisson_chunks=1)
    rs.init(); rs.step(8); rs.sync(); rs.close()
else:
    sim = ek.Simulation(ek.default_params(**over)); sim.init(); sim.step(8); sim.sync(); sim.close()
