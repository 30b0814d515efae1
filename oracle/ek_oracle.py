"""ctypes wrapper around oracle/libek_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module, and only as the checker.
The product (ek-pnp-3d_b200/) never imports anything under oracle/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libek_oracle.so")

FIELDS = ("rho", "ux", "uy", "uz", "charge", "chargen", "phi", "T", "Ex", "Ey", "Ez")

_PARAM_FIELDS = [
    ("NX", C.c_int), ("NY", C.c_int), ("NZ", C.c_int),
    ("Lx", C.c_double), ("Ly", C.c_double), ("Lz", C.c_double),
    ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
    ("uw", C.c_double), ("exf", C.c_double),
    ("CFL", C.c_double), ("dt", C.c_double), ("cs_square", C.c_double), ("rho0", C.c_double),
    ("chargeinf", C.c_double),
    ("voltage", C.c_double), ("voltage2", C.c_double),
    ("Ext", C.c_double), ("eps", C.c_double),
    ("diffu", C.c_double), ("nu", C.c_double), ("K", C.c_double),
    ("diffun", C.c_double), ("Kn", C.c_double),
    ("kB", C.c_double), ("electron", C.c_double), ("roomT", C.c_double),
    ("convertCtoCharge", C.c_double), ("PB_omega", C.c_double),
    ("D", C.c_double), ("Ra", C.c_double), ("TH", C.c_double),
    ("w0", C.c_double), ("ws", C.c_double), ("wa", C.c_double), ("wd", C.c_double),
    ("V", C.c_double), ("VC", C.c_double), ("VCn", C.c_double), ("VT", C.c_double),
    ("pb_iters", C.c_int),
]


class OracleParams(C.Structure):
    _fields_ = _PARAM_FIELDS


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, a second or two)."""
    src = os.path.join(_HERE, "ek_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libek_oracle.so"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.eko_default_params.argtypes = [C.POINTER(OracleParams)]
        L.eko_create.argtypes = [C.POINTER(OracleParams)]
        L.eko_create.restype = C.c_void_p
        for name in ("eko_destroy", "eko_initialization", "eko_init_equilibrium",
                     "eko_stream_collide_save", "eko_fast_poisson"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.eko_step.argtypes = [C.c_void_p, C.c_int]
        L.eko_field.argtypes = [C.c_void_p, C.c_int]
        L.eko_field.restype = C.POINTER(C.c_double)
        L.eko_get_populations.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        L.eko_fft1d.argtypes = [C.POINTER(C.c_double), C.c_int, C.c_int, C.c_int]
        L.eko_num_threads.restype = C.c_int
        L.eko_set_poisson_dc.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.eko_last_dc.argtypes = [C.c_void_p]
        L.eko_last_dc.restype = C.c_double
        _lib = L
    return _lib


def default_params(**over) -> OracleParams:
    """LBM.h as shipped, with overrides; lengths follow the grid when only
    NX/NY/NZ are overridden (Lx = NX*dx, Ly = NY*dy, Lz = (NZ-1)*dz)."""
    p = OracleParams()
    lib().eko_default_params(C.byref(p))
    for k, v in over.items():
        setattr(p, k, v)
    if "Lx" not in over:
        p.Lx = p.NX * p.dx
    if "Ly" not in over:
        p.Ly = p.NY * p.dy
    if "Lz" not in over:
        p.Lz = (p.NZ - 1) * p.dz
    return p


def params_dict(p: OracleParams) -> dict:
    return {name: getattr(p, name) for name, _ in _PARAM_FIELDS}


class Oracle:
    """One CPU simulation; mirrors the calls of the reference's main()."""

    def __init__(self, params: OracleParams):
        self.p = params
        self.L = lib()
        self.h = self.L.eko_create(C.byref(params))
        self.shape = (params.NZ, params.NY, params.NX)
        self.N = params.NX * params.NY * params.NZ

    def close(self):
        if self.h:
            self.L.eko_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def field(self, name: str) -> np.ndarray:
        """View (not a copy) of one macroscopic array, shape (NZ, NY, NX)."""
        ptr = self.L.eko_field(self.h, FIELDS.index(name))
        return np.ctypeslib.as_array(ptr, shape=(self.N,)).reshape(self.shape)

    def fields(self) -> dict:
        return {n: self.field(n).copy() for n in FIELDS}

    def set_fields(self, d: dict):
        for n, a in d.items():
            self.field(n)[...] = np.asarray(a, dtype=np.float64).reshape(self.shape)

    def initialization(self):
        self.L.eko_initialization(self.h)

    def init_equilibrium(self):
        self.L.eko_init_equilibrium(self.h)

    def step(self, n: int = 1):
        self.L.eko_step(self.h, int(n))

    def stream_collide_save(self):
        self.L.eko_stream_collide_save(self.h)

    def fast_poisson(self):
        self.L.eko_fast_poisson(self.h)

    def set_poisson_dc(self, mode: int, ghat0: float = 0.0):
        self.L.eko_set_poisson_dc(self.h, int(mode), float(ghat0))

    def last_dc(self) -> float:
        return self.L.eko_last_dc(self.h)

    def populations(self, s: int) -> np.ndarray:
        out = np.empty(27 * self.N, dtype=np.float64)
        self.L.eko_get_populations(self.h, s, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out.reshape((27,) + self.shape)


def perturb_fields(f: dict, amp: float = 0.05) -> dict:
    """SURVEY.md 8(d) parity perturbation: multiply c+, c-, T by
    1 + amp*sin(2 pi x/NX)*cos(2 pi y/NY)*sin(pi z/(NZ-1))."""
    NZ, NY, NX = f["rho"].shape
    z, y, x = np.meshgrid(np.arange(NZ), np.arange(NY), np.arange(NX), indexing="ij")
    g = 1.0 + amp * np.sin(2 * np.pi * x / NX) * np.cos(2 * np.pi * y / NY) * np.sin(np.pi * z / (NZ - 1))
    out = dict(f)
    for k in ("charge", "chargen", "T"):
        out[k] = f[k] * g
    return out
