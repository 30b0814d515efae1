#!/usr/bin/env python3
"""Build the reference's own CUDA code into oracle/_ref/ (TEST INFRASTRUCTURE).

The reference (gyf135/EK-PNP-3D, mounted read-only at /root/reference) has no
build system and no runtime parameters: every input is a compile-time constant
in LBM.h.  This recipe therefore

  * regenerates LBM.h per case in a scratch directory by rewriting only the
    constant initialisers listed in CASES (regex on the declaration lines),
  * compiles oracle/ref_driver.cu, which #includes that header and then the
    reference's LBM.cu and poisson.cu from where they lie (-I /root/reference),
  * writes nothing but binaries and a manifest into oracle/_ref/.

No reference source is copied into the repository.  /root/reference does not
exist on the GPU box, so this runs in the build container only (it is called
from __graft_entry__.build()); the binaries travel with the snapshot.

Also built: ek_ref_stock = the UNMODIFIED main.cu + seconds.cpp, for the
"as shipped, I/O included" timing of config C1, and the LINK-LEVEL DROP-IN
PROOF (MAIN_CASES): the reference's real main.cu compiled twice per case --

  ek_ref_main_<case>     main.cu + the reference's own LBM.cu / poisson.cu,
  ek_main_linked_<case>  the same main.cu, but its #include "LBM.cu" and
                         #include "poisson.cu" see generated files WITHOUT the
                         hot path (LBM.cu minus :68-2416 = initialization,
                         init_equilibrium, stream_collide_save and their
                         kernels; poisson.cu minus :28-204), linked against
                         ek-pnp-3d_b200/libek_b200_shim.so instead.

main.cu's quoted includes resolve relative to its own directory first, so the
build stages a byte-identical copy of main.cu next to the generated files in a
scratch directory that is deleted afterwards; nothing of it enters the repo.
"""
from __future__ import annotations

import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("EK_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")
ARCH = ["-gencode", "arch=compute_100,code=sm_100"]

DX = 1.0e-6 / 100.0

# name -> overrides of LBM.h symbols.  Lx/Ly/Lz follow the grid (App. B).
CASES = {
    # C1: the shipped case (LBM.h untouched except nothing) -- 50x8x51, nThreads 10
    "c1": dict(NX=50, NY=8, NZ=51, nThreads=10),
    # small parity grids (perturbation applied at run time by the driver)
    "g1": dict(NX=16, NY=8, NZ=13, nThreads=16),
    "g2": dict(NX=32, NY=16, NZ=17, nThreads=32),
    # NE = 32: the reference's forward cuFFT leaves an exactly zero (0,0,0) coefficient
    "g4": dict(NX=16, NY=8, NZ=17, nThreads=16),
    "g2_nofmad": dict(NX=32, NY=16, NZ=17, nThreads=32, _nvcc=["-fmad=false"]),
    # moving top wall + pressure drive + asymmetric zeta potentials
    "g3": dict(NX=24, NY=8, NZ=11, nThreads=24, uw_host=1.0e-4, exf_host=2.0e6,
               voltage2=-2.5e-3),
    # C2: isothermal electro-osmotic slit flow 128x64x64 (TH = 0 -> T == 0)
    "c2": dict(NX=128, NY=64, NZ=64, nThreads=128, TH=0.0),
    # C3: EK-PNP + temperature coupling 256^3.  With the shipped c_inf = 0.01 and
    # dz = 1e-8 the reference's own Poisson-Boltzmann start-up (LBM.cu:89-106)
    # overflows to NaN once NZ >~ 200 (the under-relaxed fixed point iteration
    # diverges for channels wider than ~20 Debye lengths); c_inf = 0.002 keeps the
    # 2.55 um channel inside the convergent range.  Everything else as shipped.
    "c3": dict(NX=256, NY=256, NZ=256, nThreads=128, chargeinf=0.002),
    "c3_t64": dict(NX=256, NY=256, NZ=256, nThreads=64, chargeinf=0.002),
    "c3_t256": dict(NX=256, NY=256, NZ=256, nThreads=256, chargeinf=0.002),
}

# symbol -> (regex matching "<decl> = <value>;", formatter)
_DECL = {
    "nThreads": r"(const int nThreads\s*=\s*)([^;]+)(;)",
    "NX": r"(const unsigned int NX\s*=\s*)([^;]+)(;)",
    "NY": r"(const unsigned int NY\s*=\s*)([^;]+)(;)",
    "NZ": r"(const unsigned int NZ\s*=\s*)([^;]+)(;)",
    "Lx": r"(__constant__ double Lx\s*=\s*)([^;]+)(;)",
    "Ly": r"(__constant__ double Ly\s*=\s*)([^;]+)(;)",
    "Lz": r"(__constant__ double Lz\s*=\s*)([^;]+)(;)",
    "uw_host": r"(double uw_host\s*=\s*)([^;]+)(;)",
    "exf_host": r"(double exf_host\s*=\s*)([^;]+)(;)",
    "voltage": r"(__constant__ double voltage\s*=\s*)([^;]+)(;)",
    "voltage2": r"(__constant__ double voltage2\s*=\s*)([^;]+)(;)",
    "Ext": r"(__constant__ double Ext\s*=\s*)([^;]+)(;)",
    "chargeinf": r"(__constant__ double chargeinf\s*=\s*)([^;]+)(;)",
    "TH": r"(__device__ double TH\s*=\s*)([^;]+)(;)",
    "Ra": r"(__device__ double Ra\s*=\s*)([^;]+)(;)",
    "NSTEPS": r"(const unsigned int NSTEPS\s*=\s*)([^;]+)(;)",
}

# the reference's real main() (dumps, diagnostics, performance block and all): as shipped, and on a grid
# whose NE = 32 is a power of two, where the reference's (0,0,0) Poisson coefficient is exactly zero
#  so that the dump files of the two builds can be compared digit by digit
MAIN_CASES = {
    "c1": dict(NX=50, NY=8, NZ=51, nThreads=10),
    "g4": dict(NX=16, NY=8, NZ=17, nThreads=16, NSTEPS=200),
}
INT_SYMS = ("nThreads", "NX", "NY", "NZ", "NSTEPS")


def case_params(name: str, table=None) -> dict:
    """Full symbol->value map written into the generated header for a case."""
    c = {k: v for k, v in (table or CASES)[name].items() if not k.startswith("_")}
    c.setdefault("Lx", c["NX"] * DX)
    c.setdefault("Ly", c["NY"] * DX)
    c.setdefault("Lz", (c["NZ"] - 1) * DX)
    return c


def patched_header(name: str, table=None) -> str:
    with open(os.path.join(REF, "LBM.h")) as f:
        text = f.read()
    for sym, val in case_params(name, table).items():
        pat = _DECL[sym]
        lit = str(int(val)) if sym in INT_SYMS else repr(float(val))
        text, n = re.subn(pat, lambda m: m.group(1) + lit + m.group(3), text, count=1)
        if n != 1:
            raise RuntimeError(f"could not rewrite {sym} in LBM.h")
    return text


def build_case(name: str, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, f"ek_ref_{name}")
    with tempfile.TemporaryDirectory(prefix="ekref_") as tmp:
        with open(os.path.join(tmp, "LBM.h"), "w") as f:
            f.write(patched_header(name))
        cmd = ["nvcc", "-O3", "-w", *ARCH, *CASES[name].get("_nvcc", []),
               "-I", tmp, "-I", REF, os.path.join(HERE, "ref_driver.cu"),
               "-lcufft", "-o", exe]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return exe


def build_stock() -> str:
    """The reference exactly as shipped (main.cu + seconds.cpp)."""
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, "ek_ref_stock")
    cmd = ["nvcc", "-O3", "-w", *ARCH, os.path.join(REF, "main.cu"),
           os.path.join(REF, "seconds.cpp"), "-lcufft", "-o", exe]
    subprocess.check_call(cmd)
    return exe


def shim_env(name: str) -> str:
    """EK_SHIM_PARAMS for ek_main_linked_<name>: the constants of the generated LBM.h"""
    c = case_params(name, MAIN_CASES)
    return ",".join(f"{k}={v!r}" for k, v in c.items() if k not in ("nThreads", "NSTEPS"))


def _without_lines(path: str, first: int, last: int) -> str:
    """the text of a reference file minus the 1-based line range [first, last]"""
    with open(path) as f:
        lines = f.readlines()
    return "".join(lines[:first - 1] + lines[last:])


def build_main(name: str, linked: bool, verbose: bool = False) -> str:
    """The reference's real main.cu on the constants of MAIN_CASES[name]; linked=True swaps its hot
    path (LBM.cu:68-2416, poisson.cu:28-204) for libek_b200_shim.so at link time."""
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, (f"ek_main_linked_{name}" if linked else f"ek_ref_main_{name}"))
    pkg = os.path.join(os.path.dirname(HERE), "ek-pnp-3d_b200")
    with tempfile.TemporaryDirectory(prefix="ekmain_") as tmp:
        with open(os.path.join(tmp, "LBM.h"), "w") as f:
            f.write(patched_header(name, MAIN_CASES))
        shutil.copy(os.path.join(REF, "main.cu"), os.path.join(tmp, "main.cu"))   # scratch only (see docstring)
        cmd = ["nvcc", "-O3", "-w", *ARCH, "-I", tmp, "-I", REF, os.path.join(tmp, "main.cu"),
               os.path.join(REF, "seconds.cpp"), "-lcufft", "-o", exe]
        if linked:
            with open(os.path.join(tmp, "LBM.cu"), "w") as f:
                f.write(_without_lines(os.path.join(REF, "LBM.cu"), 68, 2416))
            with open(os.path.join(tmp, "poisson.cu"), "w") as f:
                f.write(_without_lines(os.path.join(REF, "poisson.cu"), 28, 10 ** 9))
            cmd += ["-L", pkg, "-lek_b200_shim", "-lek_b200",
                    "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../ek-pnp-3d_b200"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return exe


def main(argv):
    if not os.path.isdir(REF):
        print(f"{REF} not present: nothing to build (prebuilt binaries are used)")
        return 0
    names = argv[1:] or list(CASES)
    manifest = {}
    if not argv[1:]:
        build_stock()
        manifest["stock"] = {"binary": "ek_ref_stock", "note": "unmodified main.cu"}
    if not argv[1:] or "main" in argv[1:]:
        names = [n for n in names if n != "main"]
        for n in MAIN_CASES:
            build_main(n, linked=False, verbose=True)
            build_main(n, linked=True, verbose=True)
            manifest[f"main_{n}"] = {"binaries": [f"ek_ref_main_{n}", f"ek_main_linked_{n}"],
                                     "params": case_params(n, MAIN_CASES), "shim_env": shim_env(n)}
    for n in names:
        build_case(n, verbose=True)
        manifest[n] = {"binary": f"ek_ref_{n}", "params": case_params(n),
                       "nvcc": CASES[n].get("_nvcc", [])}
    mpath = os.path.join(OUT, "manifest.json")
    old = {}
    if os.path.exists(mpath):
        with open(mpath) as f:
            old = json.load(f)
    old.update(manifest)
    with open(mpath, "w") as f:
        json.dump(old, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
