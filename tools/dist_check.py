#!/usr/bin/env python3
"""Parity of the one-process-per-GPU path (torch.distributed/NCCL) on real GPUs: every rank runs its
x-slab of a small perturbed problem, rank 0 also runs the single-domain simulation and compares.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py
"""
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ek = importlib.import_module("ek-pnp-3d_b200")
    slab = importlib.import_module("ek-pnp-3d_b200.slab")
    from oracle import ek_oracle as eo
    over = dict(NX=64 * world, NY=12, NZ=21, pb_iters=30, exf=1.0e6, uw=1.0e-4, voltage2=-3.0e-3)
    # the same perturbed start on every rank (deterministic)
    o = eo.Oracle(eo.default_params(**over))
    o.set_poisson_dc(0)
    o.initialization()
    init = eo.perturb_fields(o.fields(), 0.05)
    o.close()
    out = {}
    for transport in ("nccl", "dma"):
        grp = slab.SlabGroup(ek, ek.default_params(**over), slab.DistComm(dist), device=local)
        used = grp.set_transport(transport)
        grp.set_fields(init)
        grp.init_equilibrium()
        grp.step(4)
        grp.step(3)
        mine = grp.slabs[0].sim.fields()
        grp.close()
        # gather the slabs on rank 0
        parts = {}
        for k in ek.FIELDS:
            t = torch.from_numpy(np.ascontiguousarray(mine[k])).cuda()
            every = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
            dist.gather(t, every, dst=0)
            if rank == 0:
                parts[k] = np.concatenate([e.cpu().numpy() for e in every], axis=2)
        out[used] = parts
    if rank == 0:
        sim = ek.Simulation(ek.default_params(**over), device=local)
        sim.set_fields(init)
        sim.init_equilibrium()
        sim.step(4)
        sim.step(3)
        want = sim.fields()
        sim.close()
        rep = {"world": world, "grid": [over["NX"], over["NY"], over["NZ"]], "steps": 7}
        ok = True
        for used, parts in out.items():
            errs = {}
            for grp_name, names in (("rho", ("rho",)), ("u", ("ux", "uy", "uz")), ("charge", ("charge", "chargen")),
                                    ("phi", ("phi",)), ("T", ("T",)), ("E", ("Ex", "Ey", "Ez"))):
                scale = max(np.abs(want[n]).max() for n in names) or 1.0
                errs[grp_name] = float(max(np.abs(parts[n] - want[n]).max() for n in names) / scale)
            rep[used] = errs
            ok = ok and all(v <= (1e-7 if g == "u" else 1e-12) for g, v in errs.items())
        if len(out) == 2:
            a, b = list(out.values())
            rep["transports_bitwise_identical"] = bool(all(np.array_equal(a[k], b[k]) for k in ek.FIELDS))
            ok = ok and rep["transports_bitwise_identical"]
        rep["ok"] = bool(ok)
        print(json.dumps(rep), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
