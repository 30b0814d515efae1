#!/usr/bin/env python3
"""Parity of the native one-process-per-GPU driver (ek_rank.cu, NCCL from C++) on real GPUs: every
rank runs its x-slab of a perturbed problem, rank 0 also runs the single-domain simulation of the
same library and compares all 11 fields.  Cases: a generic one and a C4-shaped one (NX = 1024 split
over the ranks, small NY/NZ, pressure + electro-driven), plus the library's own start-up.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/rank_check.py
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
from tests import util  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ek = importlib.import_module("ek-pnp-3d_b200")
    from oracle import ek_oracle as eo
    nbytes = ek.load_library().ek_rank_nccl_id_bytes()

    def bcast(raw):
        t = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    def gather(mine):
        parts = {}
        for k in ek.FIELDS:
            t = torch.from_numpy(np.ascontiguousarray(mine[k])).cuda()
            every = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
            dist.gather(t, every, dst=0)
            if rank == 0:
                parts[k] = np.concatenate([e.cpu().numpy() for e in every], axis=2)
        return parts

    report = {"world": world, "nccl": ek.load_library().ek_rank_nccl_version(), "cases": {}}
    ok = True
    cases = {
        "generic": dict(NX=64 * world, NY=12, NZ=21, pb_iters=30, exf=1.0e6, uw=1.0e-4, voltage2=-3.0e-3),
        "three_tiles_per_rank": dict(NX=96 * world, NY=6, NZ=21, pb_iters=30, exf=1.0e6, voltage2=-3.0e-3),
        "c4_shaped": dict(NX=1024, NY=8, NZ=37, pb_iters=30, exf=2.0e6, chargeinf=0.002),
    }
    for name, over in cases.items():
        if over["NX"] % (2 * world):
            continue
        o = eo.Oracle(eo.default_params(**over))      # the same perturbed start on every rank (deterministic)
        o.set_poisson_dc(0)
        o.initialization()
        init = eo.perturb_fields(o.fields(), 0.05)
        o.close()
        w = over["NX"] // world
        res = {}
        for overlap in (True, False):
            rs = ek.RankSimulation(ek.default_params(**over), local, rank, world, bcast, poisson_chunks=3)
            rs.set_pipeline(overlap, overlap, boundary_first=overlap)
            rs.set_fields({k: np.ascontiguousarray(v[:, :, rank * w:(rank + 1) * w]) for k, v in init.items()})
            rs.init_equilibrium()
            rs.step(4)
            rs.step(3)
            res[overlap] = gather(rs.fields())
            rs.close()
        # the library's own start-up on the ranks
        rs = ek.RankSimulation(ek.default_params(**over), local, rank, world, bcast)
        rs.init()
        rs.step(2)
        started = gather(rs.fields())
        rs.close()
        if rank == 0:
            sim = ek.Simulation(ek.default_params(**over), device=local)
            sim.set_fields(init)
            sim.init_equilibrium()
            sim.step(4)
            sim.step(3)
            want = sim.fields()
            sim.close()
            sim = ek.Simulation(ek.default_params(**over), device=local)
            sim.init()
            sim.step(2)
            want_started = sim.fields()
            sim.close()
            err = util.field_errors(res[True], want)
            err_start = util.field_errors(started, want_started)
            same = all(np.array_equal(res[True][k], res[False][k]) for k in ek.FIELDS)
            good = same
            for e in (err, err_start):
                good = good and all(e[g] <= 1e-12 for g in ("rho", "charge", "chargen", "phi", "T", "E"))
                good = good and e["u_abs"] <= 1e-12 * e["u_scale"] + 16 * e["u_ulp"]
            report["cases"][name] = {"grid": [over["NX"], over["NY"], over["NZ"]], "steps": 7,
                                     "ranks_vs_single_domain": {k: float(v) for k, v in err.items()},
                                     "startup_then_2_steps": {k: float(v) for k, v in err_start.items()},
                                     "overlapped_equals_sequential_bitwise": bool(same), "ok": bool(good)}
            ok = ok and good
    if rank == 0:
        report["ok"] = bool(ok)
        print(json.dumps(report), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
