import importlib, sys, time, os
sys.path.insert(0, '/root/repo')
import torch, numpy as np
ek = importlib.import_module("ek-pnp-3d_b200")
p = ek.default_params(NX=256, NY=256, NZ=256, pb_iters=20, chargeinf=0.002)
sim = ek.Simulation(p); sim.init(); sim.step(4); sim.sync()
host = {n: torch.empty((256,256,256), dtype=torch.float64, pin_memory=True).numpy() for n in ek.FIELDS}
out = {n: torch.empty((256,256,256), dtype=torch.float64, pin_memory=True).numpy() for n in ek.FIELDS}
for n in ek.FIELDS: sim.field(n, out=host[n])
for rep in range(3):
    t0=time.perf_counter(); sim.run_from_host(host, 20, out); sim.sync(); print("job", time.perf_counter()-t0, flush=True)
t0=time.perf_counter(); sim.set_fields(host); sim.sync(); t1=time.perf_counter(); print("set_fields", t1-t0)
sim.init_equilibrium(); sim.sync(); t2=time.perf_counter(); print("init_eq", t2-t1)
sim.step(20); sim.sync(); t3=time.perf_counter(); print("steps", t3-t2)
for n in ek.FIELDS: sim.field(n, out=out[n])
sim.sync(); print("get", time.perf_counter()-t3)
