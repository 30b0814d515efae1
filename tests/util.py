"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import importlib
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

FIELDS = ("rho", "ux", "uy", "uz", "charge", "chargen", "phi", "T", "Ex", "Ey", "Ez")
# fields that share one scale (components of one vector)
GROUPS = {"rho": ("rho",), "u": ("ux", "uy", "uz"), "charge": ("charge",), "chargen": ("chargen",),
          "phi": ("phi",), "T": ("T",), "E": ("Ex", "Ey", "Ez")}


def ek_module():
    return importlib.import_module("ek-pnp-3d_b200")


def case_overrides(case: str) -> dict:
    """LBM.h overrides of a reference build case (oracle/build_ref.py CASES)."""
    from oracle.build_ref import case_params
    o = dict(case_params(case))
    o.pop("nThreads", None)
    ren = {"uw_host": "uw", "exf_host": "exf"}
    return {ren.get(k, k): v for k, v in o.items()}


def product_params(case_or_over):
    over = case_overrides(case_or_over) if isinstance(case_or_over, str) else dict(case_or_over)
    return ek_module().default_params(**over)


def oracle_params(case_or_over):
    from oracle import ek_oracle as eo
    over = case_overrides(case_or_over) if isinstance(case_or_over, str) else dict(case_or_over)
    return eo.default_params(**over)


def have_ref(case: str) -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"ek_ref_{case}"))


def read_fields(path: str, shape) -> dict:
    n = int(np.prod(shape))
    raw = np.fromfile(path, dtype=np.float64)
    assert raw.size == 11 * n, (raw.size, n)
    return {name: raw[i * n:(i + 1) * n].reshape(shape).copy() for i, name in enumerate(FIELDS)}


def write_fields(path: str, fields: dict):
    with open(path, "wb") as f:
        for name in FIELDS:
            np.ascontiguousarray(fields[name], dtype=np.float64).tofile(f)


def read_pops(path: str, shape) -> np.ndarray:
    n = int(np.prod(shape))
    raw = np.fromfile(path, dtype=np.float64)
    assert raw.size == 4 * 27 * n
    return raw.reshape((4, 27) + tuple(shape))


def run_ref(case: str, steps: int, perturb: float = 0.0, init_fields: dict | None = None, pops: bool = False,
            extra=(), dc: bool = False, dump_init: bool = True):
    """Run the reference's own CUDA build (oracle/_ref) and return
    (init_fields, final_fields, pops or None, info json); info["dc"] holds the
    per-step forward DC coefficients when dc=True."""
    from oracle.build_ref import case_params
    cp = case_params(case)
    shape = (cp["NZ"], cp["NY"], cp["NX"])
    exe = os.path.join(REF_DIR, f"ek_ref_{case}")
    with tempfile.TemporaryDirectory(prefix="ekref_run_") as tmp:
        cmd = [exe, "--steps", str(steps), "--dump-final", os.path.join(tmp, "final.bin")]
        if dump_init:
            cmd += ["--dump-init", os.path.join(tmp, "init.bin")]
        if perturb:
            cmd += ["--perturb", repr(perturb)]
        if init_fields is not None:
            write_fields(os.path.join(tmp, "load.bin"), init_fields)
            cmd += ["--load-init", os.path.join(tmp, "load.bin")]
        if pops:
            cmd += ["--dump-pops", os.path.join(tmp, "pops.bin")]
        if dc:
            cmd += ["--dump-dc", os.path.join(tmp, "dc.bin")]
        cmd += list(extra)
        out = subprocess.run(cmd, check=True, capture_output=True, text=True, cwd=tmp).stdout
        info = json.loads(out.strip().splitlines()[-1])
        if dc:
            info["dc"] = np.fromfile(os.path.join(tmp, "dc.bin"), dtype=np.float64)
        init = read_fields(os.path.join(tmp, "init.bin"), shape) if dump_init else init_fields
        final = read_fields(os.path.join(tmp, "final.bin"), shape)
        p = read_pops(os.path.join(tmp, "pops.bin"), shape) if pops else None
    return init, final, p, info


def u_ulp(ref: dict, cfl: float = 0.01) -> float:
    """One ulp of the largest fluid population expressed as a velocity.  u = (sum_d c_d f_d / CFL +
    F dt/2) / rho (LBM.cu:639-644) is a difference of populations of size w0*rho ~ 3e2 that cancel
    to ~1e-5 or less, so an implementation that is not bit-identical in every population (FMA
    contraction, FFT order) differs in u by a few of these, whatever max|u| is."""
    rho = float(np.abs(ref["rho"]).max())
    return float(np.spacing(8.0 / 27.0 * rho)) / (cfl * rho)


def field_errors(a: dict, b: dict) -> dict:
    """max|a-b| / max|b| per field group (components of a vector share the scale); for the velocity
    also the absolute error, its scale and the error in units of u_ulp()."""
    out = {}
    for g, names in GROUPS.items():
        scale = max(float(np.abs(b[n]).max()) for n in names)
        err = max(float(np.abs(np.asarray(a[n]) - np.asarray(b[n])).max()) for n in names)
        out[g] = err / scale if scale > 0 else err
        if g == "u":
            out["u_abs"], out["u_scale"], out["u_ulp"] = err, scale, u_ulp(b)
            out["u_ulps"] = err / out["u_ulp"]
    return out


def pop_errors(a: np.ndarray, b: np.ndarray) -> list:
    """per set: max|a-b| / max|b|"""
    res = []
    for s in range(a.shape[0]):
        scale = float(np.abs(b[s]).max())
        res.append(float(np.abs(a[s] - b[s]).max()) / (scale if scale > 0 else 1.0))
    return res
