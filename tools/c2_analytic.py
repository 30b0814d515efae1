#!/usr/bin/env python3
"""Config C2: isothermal electro-osmotic slit flow 128x64x64 run to steady state and
compared with the analytic profiles (development aid; the assertions live in
tests/test_parity_gpu.py::test_c2_slit_flow_matches_the_analytic_profiles).

  * potential: linearised Poisson-Boltzmann (Debye-Hueckel) between two walls at zeta,
        phi(z) = zeta cosh(kappa (z - H/2)) / cosh(kappa H/2),   kappa^2 = 2 F c_inf e / (eps kB T0)
  * velocity: Helmholtz-Smoluchowski with the full potential,
        u_x(z) = eps Ext (phi(z) - zeta) / (rho0 nu)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import util  # noqa: E402


def run(ek, steps, NX=128, NY=64, NZ=64):
    p = ek.default_params(NX=NX, NY=NY, NZ=NZ, TH=0.0, exf=0.0, Ext=1.0e4)
    sim = ek.Simulation(p)
    sim.init()
    hist = []
    done = 0
    for n in steps:
        sim.step(n - done)
        done = n
        f = {k: sim.field(k) for k in ("ux", "phi", "charge", "chargen", "rho", "uz")}
        hist.append((n, analyse(p, f)))
    sim.close()
    return hist


def analyse(p, f):
    NZ = p.NZ
    z = np.arange(NZ) * p.dz
    H = (NZ - 1) * p.dz
    zeta = p.voltage
    kappa = np.sqrt(2.0 * p.convertCtoCharge * p.chargeinf * p.electron / (p.eps * p.kB * p.roomT))
    phi_dh = zeta * np.cosh(kappa * (z - 0.5 * H)) / np.cosh(0.5 * kappa * H)
    phi = f["phi"][:, 0, 0]
    ux = f["ux"][:, 0, 0]
    u_hs = p.eps * p.Ext * (phi - zeta) / (p.rho0 * p.nu)
    u_scale = np.abs(u_hs).max()
    interior = slice(1, NZ - 1)
    # the reference's full-way bounce-back puts the no-slip plane half a cell inside the wall node,
    # the Dirichlet condition of the potential sits on the node itself
    phi_slip = 0.5 * (0.5 * (phi[0] + phi[1]) + 0.5 * (phi[-1] + phi[-2]))
    u_hs_half = p.eps * p.Ext * (phi - phi_slip) / (p.rho0 * p.nu)
    phi_slip_dh = zeta * np.cosh(kappa * (0.5 * p.dz - 0.5 * H)) / np.cosh(0.5 * kappa * H)
    u_dh_half = p.eps * p.Ext * (phi_dh - phi_slip_dh) / (p.rho0 * p.nu)
    return {"u_vs_hs_halfway": float(np.abs(ux[interior] - u_hs_half[interior]).max() / u_scale),
            "u_vs_dh_halfway": float(np.abs(ux[interior] - u_dh_half[interior]).max() / u_scale),
            "u_profile": [float(v) for v in ux[:6]], "u_hs_half_profile": [float(v) for v in u_hs_half[:6]],"kappa_H": float(kappa * H), "debye_cells": float(1.0 / kappa / p.dz),
            "phi_vs_dh": float(np.abs(phi - phi_dh).max() / abs(zeta)),
            "u_vs_hs": float(np.abs(ux[interior] - u_hs[interior]).max() / u_scale),
            "u_mid": float(ux[NZ // 2]), "u_hs_mid": float(u_hs[NZ // 2]),
            "uniform_xy": float(np.abs(f["ux"] - f["ux"][:, :1, :1]).max() / u_scale),
            "uz_max": float(np.abs(f["uz"]).max() / u_scale)}


if __name__ == "__main__":
    ek = util.ek_module()
    steps = [int(v) for v in sys.argv[1:]] or [2000, 5000, 10000, 15000, 20000]
    for n, a in run(ek, steps):
        print(json.dumps({"steps": n, **a}), flush=True)
