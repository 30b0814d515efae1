"""ek-pnp-3d_b200 -- B200-native coupled electrokinetic time step (host mirror).

Thin ctypes layer over ``libek_b200.so`` (include/ek_b200.h).  The Python
names mirror the reference's host functions for the hot path
(``initialization``, ``init_equilibrium``, ``stream_collide_save``,
``fast_Poisson``; LBM.h:159-176 of gyf135/EK-PNP-3D) so that tests read like
the reference's ``main()``.

There is no CPU path: importing works anywhere (the library only needs the
CUDA runtime to load), but creating a simulation without the built library or
without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EK_B200_LIB") or os.path.join(_HERE, "libek_b200.so")  # override: development A/B builds
# the same sources built with -DEK_XCHECK (slower kernel variants, literal odd-extension Poisson transform):
# what the parity tests cross-check the product against; never loaded by the product path
XCHECK_LIB_PATH = os.path.join(_HERE, "libek_b200_xcheck.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "ek_b200.h")

FIELDS = ("rho", "ux", "uy", "uz", "charge", "chargen", "phi", "T", "Ex", "Ey", "Ez")
SETS = ("fluid", "cation", "anion", "temperature")
STREAM_AA, STREAM_PUSH = 0, 1
DC_ZERO, DC_LITERAL, DC_PRESCRIBED = 0, 1, 2

_STATUS = {0: "EK_OK", 1: "EK_ERR_INVALID", 2: "EK_ERR_CUDA", 3: "EK_ERR_CUFFT", 4: "EK_ERR_STATE", 5: "EK_ERR_NOMEM"}


class EkError(RuntimeError):
    pass


class Params(C.Structure):
    """ek_params of include/ek_b200.h (the constants of LBM.h:29-125)."""
    _fields_ = [
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

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


_libs = {}


def load_library(path: str | None = None):
    """Load libek_b200.so (or another build of it); fail loudly if it has not been built."""
    path = path or LIB_PATH
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise EkError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C ek-pnp-3d_b200/csrc`. There is no CPU fallback.")
    L = C.CDLL(path)
    H = C.c_void_p
    L.ek_abi_version.restype = C.c_int
    L.ek_device_count.restype = C.c_int
    L.ek_default_params.argtypes = [C.POINTER(Params)]
    L.ek_default_params.restype = None
    L.ek_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(H)]
    for name in ("ek_destroy", "ek_init_fields", "ek_init_equilibrium", "ek_init", "ek_sync", "ek_reset_counters"):
        getattr(L, name).argtypes = [H]
    L.ek_set_fields.argtypes = [H, C.POINTER(C.c_void_p), C.c_int]
    L.ek_step.argtypes = [H, C.c_int]
    L.ek_run_from_host.argtypes = [H, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]
    L.ek_step_timed.argtypes = [H, C.c_int, C.POINTER(C.c_float)]
    L.ek_stream_collide_save.argtypes = [H, C.c_int]
    L.ek_stream_collide_save_range.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ek_stream_collide_save_part.argtypes = [H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ek_switch_stream.argtypes = [H, C.c_void_p]
    L.ek_fast_poisson.argtypes = [H, C.c_int]
    L.ek_get_field.argtypes = [H, C.c_int, C.c_void_p, C.c_int]
    L.ek_field_ptr.argtypes = [H, C.c_int, C.POINTER(C.c_void_p)]
    L.ek_get_populations.argtypes = [H, C.c_int, C.c_void_p, C.c_int]
    L.ek_set_option.argtypes = [H, C.c_char_p, C.c_longlong]
    L.ek_set_poisson_dc.argtypes = [H, C.c_int, C.c_double]
    L.ek_get_counter.argtypes = [H, C.c_char_p, C.POINTER(C.c_double)]
    L.ek_stream.argtypes = [H]
    L.ek_stream.restype = C.c_void_p
    L.ek_last_error.argtypes = [H]
    L.ek_last_error.restype = C.c_char_p
    # multi-GPU slabs
    L.ek_create_slab.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.POINTER(H)]
    L.ek_set_stream.argtypes = [H, C.c_void_p]
    L.ek_halo_doubles.argtypes = [H]
    L.ek_halo_doubles.restype = C.c_longlong
    L.ek_halo_pack.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p]
    L.ek_halo_unpack.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p]
    L.ek_phi_halo_pack.argtypes = [H, C.c_void_p, C.c_void_p]
    L.ek_phi_halo_unpack.argtypes = [H, C.c_void_p, C.c_void_p]
    L.ek_phi_halo_pack_range.argtypes = [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ek_phi_halo_unpack_range.argtypes = [H, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.ek_dq_ptr.argtypes = [H, C.POINTER(C.c_void_p)]
    L.ek_zsolve_columns.argtypes = [H, C.c_void_p, C.c_int, C.c_int]
    L.ek_poisson_finish.argtypes = [H, C.c_int]
    for name in ("ek_compute_efield", "ek_init_uniform", "ek_pbe", "ek_pbe_relax", "ek_ensure_allocated",
                 "ek_mark_fields_ready", "ek_refresh_charge_difference", "ek_row_pitch", "ek_lbm_parity"):
        getattr(L, name).argtypes = [H]
    L.ek_slab_poisson_setup.argtypes = [H, C.c_int]
    L.ek_slab_poisson_chunks.argtypes = [H]
    L.ek_slab_poisson_chunk.argtypes = [H, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]
    L.ek_slab_poisson_ipc_bytes.argtypes = []
    L.ek_slab_poisson_set_dma.argtypes = [H, C.c_int]
    L.ek_slab_poisson_ipc_export.argtypes = [H, C.c_void_p]
    L.ek_slab_poisson_ipc_import.argtypes = [H, C.c_int, C.c_void_p]
    L.ek_slab_poisson_set_peer.argtypes = [H, C.c_int, C.c_void_p, C.c_void_p]
    L.ek_slab_poisson_my_buffers.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    for name in ("ek_slab_poisson_forward", "ek_slab_poisson_gather_x", "ek_slab_poisson_scatter_x",
                 "ek_slab_poisson_backward", "ek_slab_poisson_push_x", "ek_slab_poisson_push_back"):
        getattr(L, name).argtypes = [H, C.c_int]
    L.ek_slab_poisson_solve.argtypes = [H]
    L.ek_slab_poisson_enable_ghosts.argtypes = [H]
    L.ek_slab_poisson_chunk_back.argtypes = [H, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]
    L.ek_slab_poisson_scatter_xg.argtypes = [H, C.c_int]
    L.ek_slab_poisson_backward_g.argtypes = [H, C.c_int]
    L.ek_adopt_field.argtypes = [H, C.c_int, C.c_void_p]
    L.ek_wall_current.argtypes = [H, C.POINTER(C.c_double)]
    L.ek_max_uz.argtypes = [H, C.POINTER(C.c_double)]
    L.ek_save_data_tecplot.argtypes = [H, C.c_char_p, C.c_double, C.c_int, C.c_int]
    L.ek_save_data_end.argtypes = [H, C.c_char_p, C.c_double]
    L.ek_read_data.argtypes = [H, C.c_char_p, C.POINTER(C.c_double)]
    L.ek_checkpoint_save.argtypes = [H, C.c_char_p, C.c_double]
    L.ek_checkpoint_load.argtypes = [H, C.c_char_p, C.POINTER(C.c_double)]
    L.ek_set_populations.argtypes = [H, C.c_int, C.c_void_p]
    L.ek_populations_restored.argtypes = [H]
    # native single-process multi-GPU driver (ek_multi.cu)
    L.ek_multi_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(H)]
    for name in ("ek_multi_destroy", "ek_multi_init_fields", "ek_multi_init_equilibrium", "ek_multi_init",
                 "ek_multi_sync", "ek_multi_slabs"):
        getattr(L, name).argtypes = [H]
    L.ek_multi_step.argtypes = [H, C.c_int]
    L.ek_multi_set_pipeline.argtypes = [H, C.c_int]
    L.ek_multi_read_data.argtypes = [H, C.c_char_p, C.POINTER(C.c_double)]
    L.ek_multi_checkpoint_save.argtypes = [H, C.c_char_p, C.c_double]
    L.ek_multi_checkpoint_load.argtypes = [H, C.c_char_p, C.POINTER(C.c_double)]
    L.ek_multi_step_timed.argtypes = [H, C.c_int, C.POINTER(C.c_float)]
    L.ek_multi_get_field.argtypes = [H, C.c_int, C.c_void_p]
    L.ek_multi_set_fields.argtypes = [H, C.POINTER(C.c_void_p)]
    L.ek_multi_slab.argtypes = [H, C.c_int]
    L.ek_multi_slab.restype = C.c_void_p
    L.ek_multi_wall_current.argtypes = [H, C.POINTER(C.c_double)]
    L.ek_multi_max_uz.argtypes = [H, C.POINTER(C.c_double)]
    L.ek_multi_last_error.argtypes = [H]
    L.ek_multi_last_error.restype = C.c_char_p
    # one process per GPU, NCCL driven from C++ (ek_rank.cu)
    L.ek_rank_nccl_id_bytes.argtypes = []
    L.ek_rank_nccl_unique_id.argtypes = [C.c_void_p]
    L.ek_rank_nccl_version.argtypes = []
    L.ek_rank_create.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(H)]
    for name in ("ek_rank_destroy", "ek_rank_init_fields", "ek_rank_init_equilibrium", "ek_rank_init", "ek_rank_sync",
                 "ek_rank_chunks"):
        getattr(L, name).argtypes = [H]
    L.ek_rank_slab.argtypes = [H]
    L.ek_rank_slab.restype = C.c_void_p
    L.ek_rank_set_pipeline.argtypes = [H, C.c_int, C.c_int]
    L.ek_rank_set_boundary_first.argtypes = [H, C.c_int]
    L.ek_rank_step.argtypes = [H, C.c_int]
    L.ek_rank_step_timed.argtypes = [H, C.c_int, C.POINTER(C.c_float)]
    L.ek_rank_get_counter.argtypes = [H, C.c_char_p, C.POINTER(C.c_double)]
    L.ek_rank_profile.argtypes = [H, C.c_int, C.c_int, C.c_char_p, C.c_int]
    L.ek_rank_last_error.argtypes = [H]
    L.ek_rank_last_error.restype = C.c_char_p
    L.ek_is_xcheck_build.argtypes = []
    _libs[path] = L
    return L


def default_params(**over) -> Params:
    """LBM.h as shipped with overrides; Lx, Ly, Lz follow the grid unless given."""
    p = Params()
    load_library().ek_default_params(C.byref(p))
    for k, v in over.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    if "Lx" not in over:
        p.Lx = p.NX * p.dx
    if "Ly" not in over:
        p.Ly = p.NY * p.dy
    if "Lz" not in over:
        p.Lz = (p.NZ - 1) * p.dz
    return p


class Simulation:
    """One coupled EK-PNP simulation on one CUDA device."""

    def __init__(self, params: Params | None = None, device: int = -1, stream_mode: int | None = None,
                 zchunk: int | None = None, profile: bool = False, slab: tuple | None = None, xcheck: bool = False):
        """slab=(rank, nranks): this handle owns the x-slab `rank` of the global
        domain described by `params` (multi-GPU path, see slab.py).  xcheck=True (tests only)
        runs on the cross-check build libek_b200_xcheck.so."""
        self.L = load_library(XCHECK_LIB_PATH if xcheck else None)
        self.p = params if params is not None else default_params()
        self.h = C.c_void_p()
        self.device = int(device)
        if slab is None:
            st = self.L.ek_create(C.byref(self.p), int(device), C.byref(self.h))
        else:
            rank, nranks = slab
            self.global_params = self.p
            st = self.L.ek_create_slab(C.byref(self.p), int(device), int(rank), int(nranks), C.byref(self.h))
            local = Params.from_buffer_copy(self.p)
            local.NX = self.p.NX // int(nranks)
            self.p = local
        if st != 0:
            self.h = C.c_void_p()
            raise EkError(f"ek_create failed: {_STATUS.get(st, st)} (is a CUDA device present? there is no CPU path)")
        self.shape = (self.p.NZ, self.p.NY, self.p.NX)
        self.ncells = self.p.NX * self.p.NY * self.p.NZ
        self.t = 0.0
        if stream_mode is not None:
            self.set_option("stream_mode", stream_mode)
        if zchunk is not None:
            self.set_option("zchunk", zchunk)
        if profile:
            self.set_option("profile", 1)

    # -- plumbing ---------------------------------------------------------
    def _ck(self, st: int, what: str):
        if st != 0:
            msg = self.L.ek_last_error(self.h)
            raise EkError(f"{what}: {_STATUS.get(st, st)}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.ek_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_option(self, key: str, value: int):
        self._ck(self.L.ek_set_option(self.h, key.encode(), int(value)), f"ek_set_option({key})")

    def set_poisson_dc(self, mode: int, ghat0: float = 0.0):
        """DC mode of the Poisson right-hand side (see ek_b200.h)."""
        self._ck(self.L.ek_set_poisson_dc(self.h, int(mode), float(ghat0)), "ek_set_poisson_dc")

    def counter(self, key: str) -> float:
        v = C.c_double()
        self._ck(self.L.ek_get_counter(self.h, key.encode(), C.byref(v)), f"ek_get_counter({key})")
        return v.value

    def reset_counters(self):
        self._ck(self.L.ek_reset_counters(self.h), "ek_reset_counters")

    def sync(self):
        self._ck(self.L.ek_sync(self.h), "ek_sync")

    @property
    def stream(self) -> int:
        return self.L.ek_stream(self.h) or 0

    # -- the reference's call sequence --------------------------------------
    def initialization(self):
        """initialization() of the reference (LBM.cu:68-146)."""
        self._ck(self.L.ek_init_fields(self.h), "ek_init_fields")
        self.t = 0.0

    def init_equilibrium(self):
        """init_equilibrium() of the reference (LBM.cu:150-463)."""
        self._ck(self.L.ek_init_equilibrium(self.h), "ek_init_equilibrium")

    def init(self):
        self.initialization()
        self.init_equilibrium()

    def stream_collide_save(self, write_fields: bool = True):
        """stream_collide_save() of the reference (LBM.cu:465-481)."""
        self._ck(self.L.ek_stream_collide_save(self.h, int(write_fields)), "ek_stream_collide_save")

    def fast_Poisson(self, write_efield: bool = True):
        """fast_Poisson() of the reference (poisson.cu:75-103)."""
        self._ck(self.L.ek_fast_poisson(self.h, int(write_efield)), "ek_fast_poisson")

    def step(self, nsteps: int = 1):
        """nsteps iterations of the loop body main.cu:189-200."""
        self._ck(self.L.ek_step(self.h, int(nsteps)), "ek_step")
        self.t += nsteps * self.p.dt

    def run_from_host(self, fields_in: dict, nsteps: int, fields_out: dict | None = None) -> dict:
        """One whole job on host arrays (ek_run_from_host): upload the 11 arrays, init_equilibrium(), nsteps
        steps, download the 11 arrays, with the PCIe copies overlapped with the first and last LBM pass.
        fields_out: preallocated (ideally pinned) arrays to fill; returned."""
        src = (C.c_void_p * len(FIELDS))()
        dst = (C.c_void_p * len(FIELDS))()
        keep = []
        out = fields_out if fields_out is not None else {}
        for i, n in enumerate(FIELDS):
            a = np.ascontiguousarray(fields_in[n], dtype=np.float64)
            if a.size != self.ncells:
                raise ValueError(f"field {n}: expected {self.ncells} values, got {a.size}")
            keep.append(a)
            src[i] = a.ctypes.data
            if n not in out:
                out[n] = np.empty(self.shape, dtype=np.float64)
            if out[n].size != self.ncells or not out[n].flags["C_CONTIGUOUS"] or out[n].dtype != np.float64:
                raise ValueError(f"output array {n} must be a contiguous float64 array of {self.ncells} values")
            dst[i] = out[n].ctypes.data
        self._ck(self.L.ek_run_from_host(self.h, src, int(nsteps), dst), "ek_run_from_host")
        self.t += nsteps * self.p.dt
        return out

    def step_timed(self, nsteps: int) -> float:
        """step(nsteps) bracketed by CUDA events on the handle's stream; returns ms."""
        ms = C.c_float()
        self._ck(self.L.ek_step_timed(self.h, int(nsteps), C.byref(ms)), "ek_step_timed")
        self.t += nsteps * self.p.dt
        return ms.value

    # -- data ---------------------------------------------------------------
    def set_fields(self, fields: dict):
        """Upload macroscopic arrays (any subset of FIELDS), shape (NZ,NY,NX)."""
        arr = (C.c_void_p * len(FIELDS))()
        keep = []
        for i, n in enumerate(FIELDS):
            if n in fields and fields[n] is not None:
                a = np.ascontiguousarray(fields[n], dtype=np.float64)
                if a.size != self.ncells:
                    raise ValueError(f"field {n}: expected {self.ncells} values, got {a.size}")
                keep.append(a)
                arr[i] = a.ctypes.data
            else:
                arr[i] = None
        self._ck(self.L.ek_set_fields(self.h, arr, 0), "ek_set_fields")

    def set_fields_device(self, ptrs: dict):
        """Same with device pointers (ints), e.g. torch tensors' data_ptr()."""
        arr = (C.c_void_p * len(FIELDS))()
        for i, n in enumerate(FIELDS):
            arr[i] = ptrs.get(n)
        self._ck(self.L.ek_set_fields(self.h, arr, 1), "ek_set_fields")

    def field(self, name: str, out: np.ndarray | None = None) -> np.ndarray:
        a = out if out is not None else np.empty(self.shape, dtype=np.float64)
        self._ck(self.L.ek_get_field(self.h, FIELDS.index(name), a.ctypes.data, 0), f"ek_get_field({name})")
        return a

    def field_to_device(self, name: str, dev_ptr: int):
        self._ck(self.L.ek_get_field(self.h, FIELDS.index(name), C.c_void_p(dev_ptr), 1), f"ek_get_field({name})")

    def fields(self) -> dict:
        return {n: self.field(n) for n in FIELDS}

    def populations(self, s: int | str) -> np.ndarray:
        """Pre-collision populations, shape (27, NZ, NY, NX), reference layout."""
        if isinstance(s, str):
            s = SETS.index(s)
        a = np.empty((27,) + self.shape, dtype=np.float64)
        self._ck(self.L.ek_get_populations(self.h, int(s), a.ctypes.data, 0), "ek_get_populations")
        return a

    # -- diagnostics and dumps (LBM.cu:2492-2753) -----------------------------
    def current(self) -> float:
        v = C.c_double()
        self._ck(self.L.ek_wall_current(self.h, C.byref(v)), "ek_wall_current")
        return v.value

    def max_uz(self) -> float:
        v = C.c_double()
        self._ck(self.L.ek_max_uz(self.h, C.byref(v)), "ek_max_uz")
        return v.value

    def save_data_tecplot(self, path: str, time: float | None = None, append: bool = False, first: bool = True):
        self._ck(self.L.ek_save_data_tecplot(self.h, path.encode(), self.t if time is None else time,
                                             int(append), int(first)), "ek_save_data_tecplot")

    def save_data_end(self, path: str, time: float | None = None):
        self._ck(self.L.ek_save_data_end(self.h, path.encode(), self.t if time is None else time),
                 "ek_save_data_end")

    # -- restart (LBM.cu:2629-2671) and exact checkpoints ------------------------
    def read_data(self, path: str) -> float:
        """read_data() of the reference: macroscopic arrays from a save_data_end file;
        follow with init_equilibrium() as main.cu does"""
        t = C.c_double()
        self._ck(self.L.ek_read_data(self.h, path.encode(), C.byref(t)), "ek_read_data")
        self.t = t.value
        return t.value

    def checkpoint_save(self, path: str, time: float | None = None):
        self._ck(self.L.ek_checkpoint_save(self.h, path.encode(), self.t if time is None else time),
                 "ek_checkpoint_save")

    def checkpoint_load(self, path: str) -> float:
        t = C.c_double()
        self._ck(self.L.ek_checkpoint_load(self.h, path.encode(), C.byref(t)), "ek_checkpoint_load")
        self.t = t.value
        return t.value


class MultiSimulation:
    """The x-slab path driven natively from this one process (ek_multi.cu): one slab per entry of
    `devices` (entries may repeat), peer access instead of NCCL.  Mirrors Simulation."""

    def __init__(self, params: Params, devices, poisson_chunks: int = 0):
        self.L = load_library()
        self.p = params
        self.h = C.c_void_p()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        st = self.L.ek_multi_create(C.byref(self.p), len(devices), devs, int(poisson_chunks), C.byref(self.h))
        if st != 0:
            self.h = C.c_void_p()
            raise EkError(f"ek_multi_create failed: {_STATUS.get(st, st)}")
        self.shape = (self.p.NZ, self.p.NY, self.p.NX)

    def _ck(self, st: int, what: str):
        if st != 0:
            msg = self.L.ek_multi_last_error(self.h)
            raise EkError(f"{what}: {_STATUS.get(st, st)}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.L.ek_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self):
        self._ck(self.L.ek_multi_init(self.h), "ek_multi_init")

    def initialization(self):
        self._ck(self.L.ek_multi_init_fields(self.h), "ek_multi_init_fields")

    def init_equilibrium(self):
        self._ck(self.L.ek_multi_init_equilibrium(self.h), "ek_multi_init_equilibrium")

    def step(self, nsteps: int = 1):
        self._ck(self.L.ek_multi_step(self.h, int(nsteps)), "ek_multi_step")

    def step_timed(self, nsteps: int) -> float:
        ms = C.c_float()
        self._ck(self.L.ek_multi_step_timed(self.h, int(nsteps), C.byref(ms)), "ek_multi_step_timed")
        return ms.value

    def set_fields(self, fields: dict):
        arr = (C.c_void_p * len(FIELDS))()
        keep = []
        for i, n in enumerate(FIELDS):
            if n in fields and fields[n] is not None:
                a = np.ascontiguousarray(fields[n], dtype=np.float64)
                keep.append(a)
                arr[i] = a.ctypes.data
            else:
                arr[i] = None
        self._ck(self.L.ek_multi_set_fields(self.h, arr), "ek_multi_set_fields")

    def field(self, name: str) -> np.ndarray:
        a = np.empty(self.shape, dtype=np.float64)
        self._ck(self.L.ek_multi_get_field(self.h, FIELDS.index(name), a.ctypes.data), "ek_multi_get_field")
        return a

    def fields(self) -> dict:
        return {n: self.field(n) for n in FIELDS}

    def current(self) -> float:
        v = C.c_double()
        self._ck(self.L.ek_multi_wall_current(self.h, C.byref(v)), "ek_multi_wall_current")
        return v.value

    def max_uz(self) -> float:
        v = C.c_double()
        self._ck(self.L.ek_multi_max_uz(self.h, C.byref(v)), "ek_multi_max_uz")
        return v.value

    def checkpoint_save(self, path: str, time: float = 0.0):
        self._ck(self.L.ek_multi_checkpoint_save(self.h, path.encode(), float(time)), "ek_multi_checkpoint_save")

    def checkpoint_load(self, path: str) -> float:
        t = C.c_double()
        self._ck(self.L.ek_multi_checkpoint_load(self.h, path.encode(), C.byref(t)), "ek_multi_checkpoint_load")
        return t.value

    def read_data(self, path: str) -> float:
        t = C.c_double()
        self._ck(self.L.ek_multi_read_data(self.h, path.encode(), C.byref(t)), "ek_multi_read_data")
        return t.value


class RankSimulation:
    """This process's x-slab of a domain split over `nranks` processes, one GPU each (ek_rank.cu):
    halos and Poisson transposes travel over NCCL, driven from C++.  `bcast(bytes_or_None) -> bytes`
    distributes rank 0's NCCL id block to every rank (torch.distributed, MPI, ...); every method that
    mirrors a reference call (init, step, ...) is collective."""

    def __init__(self, params: Params, device: int, rank: int, nranks: int, bcast=None, poisson_chunks: int = 0,
                 zchunk: int | None = None):
        try:        # if PyTorch is around, let ITS libnccl.so.2 be the one in the process: ek_rank.cu dlopens NCCL by
            import torch  # noqa: F401  (SONAME, and a system copy loaded first would shadow the one torch needs)
        except ImportError:
            pass
        self.L = load_library()
        self.global_params = params
        self.rank, self.nranks = int(rank), int(nranks)
        n = self.L.ek_rank_nccl_id_bytes()
        buf = None
        if nranks > 1:
            raw = None
            if rank == 0:
                mine = (C.c_ubyte * n)()
                if self.L.ek_rank_nccl_unique_id(mine) != 0:
                    raise EkError("ek_rank_nccl_unique_id failed (libnccl.so.2 not loadable?)")
                raw = bytes(mine)
            raw = bcast(raw)
            buf = (C.c_ubyte * n).from_buffer_copy(raw)
        self.r = C.c_void_p()
        st = self.L.ek_rank_create(C.byref(params), int(device), self.rank, self.nranks, buf, int(poisson_chunks),
                                   C.byref(self.r))
        if st != 0:
            self.r = C.c_void_p()
            raise EkError(f"ek_rank_create failed: {_STATUS.get(st, st)}")
        # a Simulation view of the slab handle (fields, options, counters); it does not own the handle
        self.sim = Simulation.__new__(Simulation)
        self.sim.L = self.L
        self.sim.h = C.c_void_p(self.L.ek_rank_slab(self.r))
        local = Params.from_buffer_copy(params)
        local.NX = params.NX // self.nranks
        self.sim.p = local
        self.sim.device = int(device)
        self.sim.shape = (local.NZ, local.NY, local.NX)
        self.sim.ncells = local.NX * local.NY * local.NZ
        self.sim.t = 0.0
        self.sim.close = lambda: None
        self.shape = self.sim.shape
        if zchunk is not None:
            raise EkError("zchunk of a rank is fixed at creation (ek_slab_poisson_setup)")

    def _ck(self, st: int, what: str):
        if st != 0:
            msg = self.L.ek_rank_last_error(self.r)
            raise EkError(f"{what}: {_STATUS.get(st, st)}: {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "r", None) and self.r.value:
            self.sim.h = C.c_void_p()
            self.L.ek_rank_destroy(self.r)
            self.r = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_pipeline(self, overlap: bool = True, overlap_back: bool = True, boundary_first: bool | None = None):
        self._ck(self.L.ek_rank_set_pipeline(self.r, int(overlap), int(overlap_back)), "ek_rank_set_pipeline")
        if boundary_first is not None:
            self._ck(self.L.ek_rank_set_boundary_first(self.r, int(boundary_first)), "ek_rank_set_boundary_first")

    def initialization(self):
        self._ck(self.L.ek_rank_init_fields(self.r), "ek_rank_init_fields")

    def init_equilibrium(self):
        self._ck(self.L.ek_rank_init_equilibrium(self.r), "ek_rank_init_equilibrium")

    def init(self):
        self._ck(self.L.ek_rank_init(self.r), "ek_rank_init")

    def step(self, nsteps: int = 1):
        self._ck(self.L.ek_rank_step(self.r, int(nsteps)), "ek_rank_step")

    def step_timed(self, nsteps: int) -> float:
        ms = C.c_float()
        self._ck(self.L.ek_rank_step_timed(self.r, int(nsteps), C.byref(ms)), "ek_rank_step_timed")
        return ms.value

    def sync(self):
        self._ck(self.L.ek_rank_sync(self.r), "ek_rank_sync")

    def counter(self, key: str) -> float:
        v = C.c_double()
        self._ck(self.L.ek_rank_get_counter(self.r, key.encode(), C.byref(v)), f"ek_rank_get_counter({key})")
        return v.value

    def chunks(self) -> int:
        return self.L.ek_rank_chunks(self.r)

    def profile(self, nsteps: int, sequential: bool) -> dict:
        """ms per step and phase (ek_rank_profile); collective"""
        import json
        buf = C.create_string_buffer(4096)
        self._ck(self.L.ek_rank_profile(self.r, int(nsteps), int(sequential), buf, 4096), "ek_rank_profile")
        return json.loads(buf.value.decode())

    # this rank's columns of the macroscopic arrays, shape (NZ, NY, NX / nranks)
    def set_fields(self, fields_local: dict):
        self.sim.set_fields(fields_local)

    def fields(self) -> dict:
        return self.sim.fields()

    def field(self, name: str, out=None):
        return self.sim.field(name, out=out)
