#!/usr/bin/env python3
"""Native single-process multi-GPU driver (ek_multi.cu): correctness against the same slabs on one
device and timing of the C5-family workload (development aid; bench.py uses one process per GPU).

usage: multi_bench.py NGPUS [steps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402


def main():
    ek = util.ek_module()
    n = int(sys.argv[1])
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    # (1) the same small problem with the slabs spread over n devices and all on device 0
    over = dict(NX=64 * n, NY=16, NZ=21, pb_iters=30, exf=1.0e6)
    res = []
    for devs in (list(range(n)), [0] * n):
        m = ek.MultiSimulation(ek.default_params(**over), devs)
        m.init()
        m.step(3)
        m.step(2)
        res.append(m.fields())
        m.close()
    same = all(np.array_equal(res[0][k], res[1][k]) for k in util.FIELDS)
    print(json.dumps({"check": "slabs on %d devices == slabs on one device (bitwise)" % n, "ok": bool(same)}), flush=True)
    # (2) timing: 1024 columns x 512 x 256 per GPU (N = 8 is config C5)
    p = ek.default_params(NX=1024 * n, NY=512, NZ=256, pb_iters=10, chargeinf=0.002, exf=2.0e6)
    m = ek.MultiSimulation(p, list(range(n)))
    m.init()
    cells = p.NX * p.NY * p.NZ
    for pipeline in (1, 0):
        m.L.ek_multi_set_pipeline(m.h, pipeline)
        m.step(4)
        ms = m.step_timed(steps)
        print(json.dumps({"driver": "ek_multi (one process, peer copies, events)", "pipeline": pipeline, "n_gpus": n,
                          "grid": [p.NX, p.NY, p.NZ], "steps": steps, "ms_per_step": round(ms / steps, 3),
                          "mlups": round(cells * steps / (ms * 1e-3) / 1e6, 1)}), flush=True)
    m.close()
    assert same


if __name__ == "__main__":
    main()
