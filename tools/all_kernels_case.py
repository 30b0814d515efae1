#!/usr/bin/env python3
"""Small runs of every kernel family (a smoke run; written for compute-sanitizer memcheck, which is
closed on this GPU pool):
single domain (general + lean path, both A-A parities, field output, push mode), Poisson paths,
start-up, x-slabs through slab.py (LocalComm, all transports) and through ek_multi."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ek = importlib.import_module("ek-pnp-3d_b200")
slab = importlib.import_module("ek-pnp-3d_b200.slab")

for mode in (ek.STREAM_AA, ek.STREAM_PUSH):
    for kernel in (0, 3):
        sim = ek.Simulation(ek.default_params(NX=40, NY=5, NZ=13, pb_iters=5, uw=1e-4), stream_mode=mode)
        sim.set_option("kernel", kernel)
        sim.init()
        sim.step(3)
        sim.step(2)
        sim.fields()
        sim.populations(1)
        sim.close()
sim = ek.Simulation(ek.default_params(NX=16, NY=4, NZ=9, pb_iters=3))
sim.set_option("poisson_path", 1)
sim.init()
sim.step(2)
sim.current(); sim.max_uz()
sim.close()
for transport in ("nccl", "p2p", "dma"):
    grp = slab.SlabGroup(ek, ek.default_params(NX=96, NY=6, NZ=13, pb_iters=3), slab.LocalComm(3), zchunk=4)
    grp.set_transport(transport)
    grp.init()
    grp.step(3)
    grp.gather_fields()
    grp.close()
m = ek.MultiSimulation(ek.default_params(NX=64, NY=6, NZ=13, pb_iters=3), [0, 0])
m.init()
m.step(3)
m.fields()
m.close()
print("all kernel families ran")
