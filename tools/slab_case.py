#!/usr/bin/env python3
"""Two x-slabs of 256^3 each on ONE GPU (slab.LocalComm): the kernels of the multi-GPU path for an
ncu launch list (ncu must not wrap a multi-rank command).  usage: slab_case.py [transport]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ek = importlib.import_module("ek-pnp-3d_b200")
slab = importlib.import_module("ek-pnp-3d_b200.slab")

transport = sys.argv[1] if len(sys.argv) > 1 else "nccl"
p = ek.default_params(NX=512, NY=256, NZ=256, pb_iters=2, chargeinf=0.002, exf=2.0e6)
grp = slab.SlabGroup(ek, p, slab.LocalComm(2))
grp.set_transport(transport)
grp.init()
grp.step(4)
import torch
torch.cuda.synchronize()
grp.close()
print("slab case done")
