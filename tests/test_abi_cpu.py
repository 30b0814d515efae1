"""The C-ABI library loads and exports every symbol include/ek_b200.h declares
(no compute: this runs on the CPU-only build box)."""
import ctypes as C
import os
import re

import pytest

from tests import util


def declared_functions():
    ek = util.ek_module()
    with open(ek.HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ek_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    ek = util.ek_module()
    L = ek.load_library()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in ek_b200.h but not exported by libek_b200.so"
    assert L.ek_abi_version() == 1
    assert L.ek_is_xcheck_build() == 0
    X = ek.load_library(ek.XCHECK_LIB_PATH)       # the cross-check build exports the same ABI
    for n in names:
        assert hasattr(X, n), n
    assert X.ek_is_xcheck_build() == 1


def test_params_struct_matches_the_oracle_layout():
    ek = util.ek_module()
    from oracle import ek_oracle as eo
    assert [f[0] for f in ek.Params._fields_] == [f[0] for f in eo.OracleParams._fields_]
    assert C.sizeof(ek.Params) == C.sizeof(eo.OracleParams)
    a, b = ek.default_params(), eo.default_params()
    for name, _ in ek.Params._fields_:
        assert getattr(a, name) == getattr(b, name), name
    # LBM.h as shipped
    assert (a.NX, a.NY, a.NZ) == (50, 8, 51)
    assert a.Lx == 50 * a.dx and a.Lz == 50 * a.dz


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    ek = util.ek_module()
    with pytest.raises(ek.EkError):
        ek.Simulation(ek.default_params())


def test_invalid_arguments_are_rejected():
    ek = util.ek_module()
    L = ek.load_library()
    h = C.c_void_p()
    p = ek.default_params(NZ=3)
    assert L.ek_create(C.byref(p), 0, C.byref(h)) == 1  # EK_ERR_INVALID before any CUDA call
    assert L.ek_create(None, 0, C.byref(h)) == 1
    assert L.ek_step(None, 1) == 1
    assert L.ek_last_error(None) == b"null handle"


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may mention it."""
    pkg = os.path.join(util.ROOT, "ek-pnp-3d_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert "ek_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, fn


def test_shim_exports_the_reference_signatures():
    """libek_b200_shim.so: the four hot-path functions of the reference with their C++ signatures
    (LBM.h:159-176), the globals fast_Poisson writes as WEAK references (resolved against main.cu when it
    is linked in), and the configuration hooks"""
    import subprocess
    so = os.path.join(util.ROOT, "ek-pnp-3d_b200", "libek_b200_shim.so")
    out = subprocess.run(["nm", "-D", "-C", so], capture_output=True, text=True, check=True).stdout
    ptr = "double*"
    for sig in ("initialization(" + ", ".join([ptr] * 11) + ")",
                "init_equilibrium(" + ", ".join([ptr] * 18) + ")",
                "stream_collide_save(" + ", ".join([ptr] * 22) + ", double, double*)",
                "fast_Poisson(" + ", ".join([ptr] * 5) + ", int)"):
        assert any(l.split(" ", 2)[-1] == sig and " T " in l for l in out.splitlines()), sig
    for g in ("phi_gpu", "Ex_gpu", "Ey_gpu", "Ez_gpu"):
        assert any(l.strip().endswith(" " + g) and " w " in l for l in out.splitlines()), g
    for f in ("ek_shim_configure", "ek_shim_bind_potential", "ek_shim_handle"):
        assert f" T {f}" in out, f


def test_reference_arm_prints_a_contract_line_without_a_gpu():
    """`bench.py --impl reference` on a box where the reference's CUDA build cannot run (this container: no GPU)
    falls back to timing the CPU restatement and still prints ONE JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU box: the reference's own build runs (covered by the driver's reference arm)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         check=True, capture_output=True, text=True, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "coupled_step_mlups" and line["unit"] == "MLUPS"
    assert line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
