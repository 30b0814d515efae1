#!/usr/bin/env python3
"""bench.py -- coupled-step MLUPS of the B200-native EK-PNP step (one JSON line).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one coupled time step (main.cu:189-200 of the reference): one fused
LBM pass over the four D3Q27 population sets plus the spectral Poisson solve.
N = 1 runs config C3 (EK-PNP + temperature, 256^3); N > 1 runs the C5 family
weak-scaled, 1024*N x 512 x 256 split into x-slabs of 1024 columns, one rank per
GPU (torchrun); N = 8 is config C5 (1.07 G cells).  --workload c4 runs the
strong-scaling config C4 (1024x256x256) instead.

value  = cells * K / device time of K steps (CUDA events on the stream the
         kernels are launched on, max over ranks), state resident in HBM.
e2e    = the same metric for a whole job through the public C ABI with HOST
         buffers: upload of the 11 macroscopic arrays from pinned memory,
         init_equilibrium, K steps, download of the 11 arrays.
roofline = the fused LBM kernel: algorithmic bytes per launch / mean launch
         time (per-launch CUDA events), against MEASURED_PEAKS.json hbm_gbs.
cpu_baseline = oracle/ (C restatement, OpenMP) on a bounded sample, rank 0.

--impl reference times the reference's own CUDA build (oracle/_ref, compiled
by oracle/build_ref.py from the unmodified sources): the reference has no CPU
path, so its CUDA build on the same B200 is the baseline.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG_STEP = 1760   # bytes per cell update of the coupled step
B_ALG_LBM = 1744    # LBM kernel alone: 4*27*16 + 8 (c+ - c- out) + 8 (phi in)

# c_inf = 0.002 for the NZ = 256 grids: with the shipped 0.01 the reference's own
# Poisson-Boltzmann start-up overflows to NaN for NZ >~ 200 (oracle/build_ref.py).
WORKLOADS = {
    "c3": dict(NX=256, NY=256, NZ=256, over=dict(chargeinf=0.002),
               name="C3 EK-PNP + temperature coupling 256x256x256 (LBM.h physics as shipped, TH=1, Ra=1; "
                    "c_inf=0.002 so that the reference's PB start-up converges)"),
    "c4": dict(NX=1024, NY=256, NZ=256, over=dict(chargeinf=0.002, exf=2.0e6),
               name="C4 pressure- and electro-driven microchannel 1024x256x256, x-slabs (exf=2e6, c_inf=0.002)"),
    "c5": dict(NX=8192, NY=512, NZ=256, over=dict(chargeinf=0.002, exf=2.0e6),
               name="C5 long high-aspect-ratio microchannel 8192x512x256 (1.07 G cells), x-slabs (exf=2e6, c_inf=0.002)"),
    "c2": dict(NX=128, NY=64, NZ=64, over=dict(TH=0.0), name="C2 isothermal slit 128x64x64"),
    "c1": dict(NX=50, NY=8, NZ=51, over={}, name="C1 shipped 50x8x51"),
}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0: float, t1: float):
        """keep the samples taken inside the timed region [t0, t1]"""
        self.t0, self.t1 = t0, t1

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        inside = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.15]
        for r in inside if inside else [r for (_, r) in self.rows]:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(steps_budget_s: float = 15.0) -> dict:
    """The CPU restatement (oracle/) timed on the host cores: bounded sample of
    the same workload (C3 physics on the C2-sized grid)."""
    from oracle import ek_oracle as eo
    NX, NY, NZ = 128, 64, 64
    p = eo.default_params(NX=NX, NY=NY, NZ=NZ, pb_iters=3)
    o = eo.Oracle(p)
    o.initialization()
    o.init_equilibrium()
    o.step(1)
    t0 = time.time()
    n = 0
    while True:
        o.step(2)
        n += 2
        if time.time() - t0 > steps_budget_s or n >= 200:
            break
    dt = time.time() - t0
    o.close()
    cores = eo.lib().eko_num_threads()
    return {"value": round(NX * NY * NZ * n / dt / 1e6, 3), "unit": "MLUPS", "cores": cores, "kind": "port",
            "sample": f"oracle/ek_oracle.c (OpenMP, {cores} threads), C3 physics on a {NX}x{NY}x{NZ} grid, "
                      f"{n} coupled steps in {dt:.1f} s"}


def ncu_traffic():
    """dram bytes per LBM launch from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "lbm_kernel_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def run_reference(args, rank: int):
    """--impl reference: the reference's own CUDA build on this GPU."""
    if rank != 0:
        return
    wl = "c3"
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    best = None
    tried = []
    for variant in ("c3", "c3_t64", "c3_t256"):
        exe = os.path.join(ref_dir, f"ek_ref_{variant}")
        if not os.path.exists(exe):
            continue
        try:
            out = subprocess.run([exe, "--steps", str(args.steps), "--warmup", str(args.warmup)], check=True,
                                 capture_output=True, text=True, timeout=900, cwd="/tmp").stdout
            info = json.loads(out.strip().splitlines()[-1])
        except Exception as e:  # noqa: BLE001
            tried.append(f"{variant}: {e}")
            continue
        tried.append(f"{variant}: nThreads={info['nThreads']} {info['mlups']:.1f} MLUPS")
        if best is None or info["mlups"] > best["mlups"]:
            best = info
        if not args.ref_all:
            break
    w = WORKLOADS[wl]
    if best is None:
        # no reference binary on this box: time the CPU restatement instead
        cb = cpu_baseline()
        line = {"impl": "reference", "metric": "coupled_step_mlups", "value": cb["value"], "unit": "MLUPS",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": w["name"], "note": "the reference's CUDA build did not run here (no binary or no GPU), CPU restatement timed instead: " + "; ".join(tried)},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    v = round(best["mlups"], 2)
    extra = {}
    if not args.no_extra and args.gpus == 1:      # once per round is enough: the N > 1 launches skip it
        import re
        import tempfile
        for name, steps2 in (("c1", 1000), ("c2", 200)):     # step-only, CUDA events around main.cu:192-198
            exe = os.path.join(ref_dir, f"ek_ref_{name}")
            try:
                out = subprocess.run([exe, "--steps", str(steps2), "--warmup", "10"], check=True, capture_output=True,
                                     text=True, timeout=600, cwd="/tmp").stdout
                info = json.loads(out.strip().splitlines()[-1])
                extra[name] = {"grid": [info["NX"], info["NY"], info["NZ"]], "steps": steps2, "nThreads": info["nThreads"],
                               "us_per_step": round(1e3 * info["ms_per_step"], 2), "mlups": round(info["mlups"], 2),
                               "init_ms": info.get("init_ms")}
            except Exception as e:  # noqa: BLE001
                extra[name] = {"error": str(e)[:200]}
        # the reference exactly as shipped (unmodified main.cu: 50x8x51, 1000 steps, dumps and diagnostics inside
        # the timed loop): its own "speed (Mlups)" line, main.cu:243,251
        try:
            with tempfile.TemporaryDirectory(prefix="ekstock_") as tmp:
                out = subprocess.run([os.path.join(ref_dir, "ek_ref_stock")], input="0\n", check=True, capture_output=True,
                                     text=True, timeout=600, cwd=tmp).stdout
            m = re.search(r"speed:\s*([0-9.eE+-]+)\s*\(Mlups\)", out)
            rt = re.search(r"clock runtime:\s*([0-9.eE+-]+)", out)
            extra["c1_as_shipped"] = {"mlups": float(m.group(1)), "clock_runtime_s": float(rt.group(1)),
                                      "note": "unmodified main.cu, I/O inside the loop (main.cu:206-222)"}
        except Exception as e:  # noqa: BLE001
            extra["c1_as_shipped"] = {"error": str(e)[:200]}
    line = {"impl": "reference", "metric": "coupled_step_mlups", "value": v, "unit": "MLUPS", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(best["ms_per_step"], 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "grid": [w["NX"], w["NY"], w["NZ"]],
                       "reference_build": "unmodified LBM.cu/poisson.cu, nvcc -O3 sm_100, nThreads=%d" % best["nThreads"],
                       "variants": tried, "init_ms": best.get("init_ms"),
                       "note": "the reference is CUDA-only and single-GPU; it runs on one B200 whatever --gpus is"},
            "cpu_baseline": {"value": v, "unit": "MLUPS", "cores": 1, "kind": "reference",
                             "sample": "reference CUDA build (no CPU path exists), one host thread, "
                                       f"{args.steps} steps of the full C3 grid on the B200"},
            "e2e": {"value": v, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "extra": extra}
    print(json.dumps(line), flush=True)


def _timed_ranks(rs, dist, torch, steps, warmup):
    """W warm-up steps, then K steps bracketed by a barrier + synchronize on both sides; device time
    (CUDA events on the rank's main stream), max over ranks"""
    rs.step(warmup)
    rs.sync()
    dist.barrier()
    torch.cuda.synchronize()
    ms_local = rs.step_timed(steps)
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([ms_local], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bench_ranks(ek, dist, torch, args, w, local_rank):
    """N > 1: one rank per GPU, every rank drives its x-slab through the native per-rank driver
    (ek_rank_* of the C ABI: NCCL send/recv halos + all-to-all transposes issued from C++)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    nbytes = ek.load_library().ek_rank_nccl_id_bytes()

    def bcast(raw):
        t = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t.copy_(torch.tensor(list(raw), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    NX, NY, NZ = w["NX"], w["NY"], w["NZ"]
    cells = NX * NY * NZ
    p = ek.default_params(NX=NX, NY=NY, NZ=NZ, pb_iters=args.pb_iters, **w["over"])
    rs = ek.RankSimulation(p, local_rank, rank, world, bcast, poisson_chunks=args.poisson_chunks)
    rs.set_pipeline(not args.no_overlap, not args.no_overlap)
    t0 = time.time()
    rs.init()
    rs.sync()
    dist.barrier()
    init_s = time.time() - t0
    sampler = ClockSampler(local_rank)
    sampler.start()
    rs.step(args.warmup)
    rs.sync()
    time.sleep(0.5)
    tw0 = time.time()
    l0 = rs.counter("kernel_launches")
    g0 = rs.counter("nccl_groups")
    ms = _timed_ranks(rs, dist, torch, args.steps, args.warmup)
    # (the warm-up steps inside _timed_ranks are counted too: per-step figures below divide by both)
    per_step = 1.0 / (args.steps + args.warmup)
    launches = (rs.counter("kernel_launches") - l0) * per_step
    groups = (rs.counter("nccl_groups") - g0) * per_step
    sampler.window(tw0, time.time())
    clocks = sampler.stop()
    mlups = cells * args.steps / (ms * 1e-3) / 1e6
    zchunk_used = int(rs.counter("zchunk"))
    K = rs.chunks()
    def _profile(r_, sequential):
        r_.sync()
        dist.barrier()          # the ranks enter together: otherwise the first phase absorbs their skew
        torch.cuda.synchronize()
        r_.step(2)              # ... and are in step with each other when the marks start
        return r_.profile(4, sequential)
    phases = {"pipelined_main_stream": _profile(rs, False), "sequential_no_overlap": _profile(rs, True)}
    nccl_version = ek.load_library().ek_rank_nccl_version()

    # ---- end to end through the C ABI with HOST buffers: every rank uploads its slab of the 11
    # macroscopic arrays from pinned memory, init_equilibrium, K steps, downloads the 11 arrays
    e2e = None
    if not args.no_e2e:
        try:
            host = {n: torch.empty(rs.shape, dtype=torch.float64, pin_memory=True).numpy() for n in ek.FIELDS}
            for n in ek.FIELDS:
                rs.field(n, out=host[n])
            rs.sync()
            dist.barrier()
            t0 = time.perf_counter()
            rs.set_fields(host)
            rs.init_equilibrium()
            rs.step(args.steps)
            for n in ek.FIELDS:
                rs.field(n, out=host[n])
            rs.sync()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            e2e = {"value": round(cells * args.steps / dt / 1e6, 2), "unit": "MLUPS",
                   "h2d_bytes_per_step": int(11 * cells * 8 / args.steps), "d2h_bytes_per_step": int(11 * cells * 8 / args.steps),
                   "job": f"every rank: upload its slab of 11 fields (pinned host) + init_equilibrium + {args.steps} steps + "
                          "download 11 fields; wall clock, max over ranks", "seconds": round(dt, 4)}
            del host
        except Exception as exc:  # noqa: BLE001  (e.g. the pinned allocation of 11 x 1 GB per rank failed)
            e2e = {"value": None, "unit": "MLUPS", "error": str(exc)[:200]}
    rs.close()

    # ---- config C4 (1024x256x256) strong scaling on the same ranks, against one GPU measured here
    c4 = None
    if not args.no_c4 and w.get("weak"):
        try:
            c4w = WORKLOADS["c4"]
            p4 = ek.default_params(NX=c4w["NX"], NY=c4w["NY"], NZ=c4w["NZ"], pb_iters=min(args.pb_iters, 101), **c4w["over"])
            cells4 = c4w["NX"] * c4w["NY"] * c4w["NZ"]
            steps4 = max(args.steps, 40)
            tried = {}
            for K4 in (1, 2, 4, 0):  # at 8.4 M cells per GPU the pipeline depth is a launch-count trade-off: report all (0 = automatic)
                r4 = ek.RankSimulation(p4, local_rank, rank, world, bcast, poisson_chunks=K4)
                r4.init()
                tried[K4] = _timed_ranks(r4, dist, torch, steps4, max(args.warmup, 10))
                if K4 == 0:
                    ph4 = {"pipelined_main_stream": _profile(r4, False), "sequential_no_overlap": _profile(r4, True)}
                r4.close()
            K4 = min(tried, key=tried.get)
            ms4 = tried[K4]
            one = torch.zeros(1, dtype=torch.float64, device="cuda")
            if rank == 0:       # the same config on ONE GPU (58 GB of populations), same start-up, same step count
                sim = ek.Simulation(p4, device=local_rank)
                sim.init()
                sim.step(10)
                one[0] = sim.step_timed(steps4)
                sim.close()
            dist.broadcast(one, 0)
            ms1 = float(one.item())
            c4 = {"workload": c4w["name"], "n_gpus": world, "steps": steps4, "poisson_chunks": K4,
                  "ms_per_step_by_chunks": {str(k): round(v / steps4, 4) for k, v in tried.items()},
                  "phase_ms_rank0_auto_chunks": ph4, "ms_per_step": round(ms4 / steps4, 4),
                  "mlups": round(cells4 * steps4 / (ms4 * 1e-3) / 1e6, 1),
                  "one_gpu_ms_per_step": round(ms1 / steps4, 4), "one_gpu_mlups": round(cells4 * steps4 / (ms1 * 1e-3) / 1e6, 1),
                  "speedup": round(ms1 / ms4, 3), "efficiency_vs_one_gpu": round(ms1 / ms4 / world, 4),
                  "note": "strong scaling: the same 67 M-cell grid on N ranks (x-slabs of 1024/N columns) and on one GPU of this box"}
        except Exception as exc:  # noqa: BLE001
            c4 = {"error": str(exc)[:300]}

    peak, peak_src = measured_peak()
    step_gbs = mlups * 1e6 * B_ALG_STEP / 1e9
    return {"metric": "coupled_step_mlups", "value": round(mlups, 2), "unit": "MLUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak" if w.get("weak") else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "grid": [NX, NY, NZ], "stream_mode": "aa", "zchunk": zchunk_used,
                       "parallelism": f"x-slabs x{world}, one process per GPU, native driver (ek_rank.cu): NCCL {nccl_version} "
                                      f"ncclSend/ncclRecv halos + grouped send/recv all-to-all Poisson transposes issued from C++, "
                                      f"{K} z-chunks, forward half and way back of the Poisson stage overlapped with the LBM launches",
                       "cells_per_gpu": cells // world, "init": "reference start-up (PB iterations) %.2f s" % init_s,
                       "l2": "per-GPU working set >> 126 MB L2",
                       "per_step": {"kernel_launches_per_rank": round(launches, 1), "nccl_groups_per_rank": round(groups, 1)},
                       "phase_ms_rank0": phases},
            "roofline": {"bound": "hbm", "achieved": round(step_gbs / world, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(step_gbs / world / peak, 4), "peak_source": peak_src,
                         "note": "whole coupled step per GPU at 1760 B/cell (kernel split is reported at N=1)",
                         "traffic": None},
            "cpu_baseline": None, "e2e": e2e, "gpu_launches": int(round(launches * args.steps)), "clocks": clocks,
            "hbm_gbs_step": round(step_gbs, 1), "extra": {"c4_strong": c4}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, help="c1|c2|c3|c4 (default c3 at N=1, c4 at N>1)")
    ap.add_argument("--stream-mode", default="aa", choices=["aa", "push"])
    ap.add_argument("--zchunk", type=int, default=None, help="z-planes per CTA of the LBM kernel (default: automatic)")
    ap.add_argument("--ref-all", action="store_true", help="reference arm: try every nThreads variant")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel", type=int, default=None, help="LBM kernel variant (ek_set_option kernel)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C1 / C2 records of the `extra` key")
    ap.add_argument("--poisson-chunks", type=int, default=0,
                    help="N>1: z-chunks of the distributed Poisson stage (0: automatic, 7 chunks of sizes 1:2:3:4:3:2:1)")
    ap.add_argument("--transport", default="nccl", choices=["nccl", "p2p", "dma"],
                    help="N>1: Poisson transposes by NCCL all-to-all or by direct peer-memory writes (CUDA IPC)")
    ap.add_argument("--driver", default="native", choices=["native", "python"],
                    help="N>1: native = ek_rank.cu (NCCL driven from C++, default); python = slab.py over torch.distributed")
    ap.add_argument("--no-c4", action="store_true", help="N>1: skip the C4 strong-scaling sub-record")
    ap.add_argument("--no-overlap", action="store_true",
                    help="N>1: do not run the Poisson forward half behind the LBM launches")
    ap.add_argument("--pb-iters", type=int, default=501,
                    help="start-up Poisson-Boltzmann iterations (501 as the reference; lower only for profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"     # keep NCCL's version banner off stdout: ONE JSON line
        # NCCL prints its banner / warnings to stdout by default: send them to stderr (stdout carries ONE JSON line)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ek = importlib.import_module("ek-pnp-3d_b200")
    # N = 1: the named 1-GPU config C3.  N > 1: weak scaling in the C5 family, 1024*N x 512 x 256
    # (134 M cells = 116 GB of populations per GPU; N = 8 is exactly BASELINE config C5).
    # --workload c4 gives the strong-scaling series of config C4 instead.
    wl = args.workload or ("c3" if world == 1 else "c5w")
    if wl == "c5w":
        WORKLOADS["c5w"] = dict(NX=1024 * world, NY=512, NZ=256, over=dict(chargeinf=0.002, exf=2.0e6), weak=True,
                                name=f"C5 family, weak scaling: {1024 * world}x512x256 x-slabs of 1024 columns per GPU "
                                     "(N=8 is config C5, 1.07 G cells; exf=2e6, c_inf=0.002)")
    w = WORKLOADS[wl]
    NX, NY, NZ = w["NX"], w["NY"], w["NZ"]
    mode = ek.STREAM_AA if args.stream_mode == "aa" else ek.STREAM_PUSH

    if world > 1:
        if args.driver == "python":
            from importlib import import_module
            slab = import_module("ek-pnp-3d_b200.slab")
            result = slab.bench_slabs(ek, dist, args, w, wl, local_rank)
        else:
            result = bench_ranks(ek, dist, torch, args, w, local_rank)
        if rank == 0:
            print(json.dumps(result), flush=True)
        dist.destroy_process_group()
        return

    cells = NX * NY * NZ
    p = ek.default_params(NX=NX, NY=NY, NZ=NZ, pb_iters=args.pb_iters, **w["over"])
    sim = ek.Simulation(p, device=local_rank, stream_mode=mode, zchunk=args.zchunk)
    if args.kernel is not None:
        sim.set_option("kernel", args.kernel)
    zchunk_used = int(sim.counter("zchunk"))
    t0 = time.time()
    sim.init()            # the reference's start-up: 501 Poisson-Boltzmann iterations + equilibrium
    sim.sync()
    init_s = time.time() - t0

    # ---- device-resident throughput ------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    sim.step(args.warmup)
    sim.sync()
    time.sleep(0.5)                     # let nvidia-smi start streaming before the timed region
    sim.step(args.warmup)
    torch.cuda.synchronize()
    tw0 = time.time()
    ms = sim.step_timed(args.steps)
    torch.cuda.synchronize()
    sampler.window(tw0, time.time())
    clocks = sampler.stop()
    mlups = cells * args.steps / (ms * 1e-3) / 1e6

    # ---- per-kernel split for the roofline (separate pass, per-launch events) ----
    sim.set_option("profile", 1)
    sim.reset_counters()
    sim.step(args.steps)
    sim.sync()
    lbm_ms = sim.counter("lbm_ms") / args.steps
    poi_ms = sim.counter("poisson_ms") / args.steps
    launches = int(sim.counter("kernel_launches"))
    sim.set_option("profile", 0)
    peak, peak_src = measured_peak()
    achieved = cells * B_ALG_LBM / (lbm_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "kernel": "ek_step_kernel (fused stream+collide, 4 sets)",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "peak_source": peak_src, "alg_bytes_per_cell": B_ALG_LBM,
                "kernel_ms": round(lbm_ms, 4), "poisson_ms": round(poi_ms, 4),
                "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                "step_achieved": round(mlups * 1e6 * B_ALG_STEP / 1e9, 1),
                "step_frac": round(mlups * 1e6 * B_ALG_STEP / 1e9 / peak, 4),
                "step_frac_of_8TBps": round(mlups * 1e6 * B_ALG_STEP / 1e9 / 8000.0, 4)}

    # ---- end to end through the C ABI with host buffers ------------------
    e2e = None
    if not args.no_e2e:
        host = {n: torch.empty((NZ, NY, NX), dtype=torch.float64, pin_memory=True).numpy() for n in ek.FIELDS}
        for n in ek.FIELDS:
            sim.field(n, out=host[n])
        outb = {n: torch.empty((NZ, NY, NX), dtype=torch.float64, pin_memory=True).numpy() for n in ek.FIELDS}
        sim.run_from_host(host, 2, outb)      # untimed warm-up of the call (copy stream, events, first DMA to `outb`)
        sim.sync()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # one call of the public C ABI (ek_run_from_host): H2D of the 11 arrays from pinned memory,
        # init_equilibrium, K steps, D2H of the 11 arrays; copies pipelined against the first / last LBM pass
        sim.run_from_host(host, args.steps, outb)
        sim.sync()
        dt = time.perf_counter() - t0
        # the same job as separate calls (upload, init_equilibrium, steps, 11 downloads), for comparison
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        sim.set_fields(host)
        sim.init_equilibrium()
        sim.step(args.steps)
        for n in ek.FIELDS:
            sim.field(n, out=outb[n])
        sim.sync()
        dt_plain = time.perf_counter() - t1
        e2e = {"value": round(cells * args.steps / dt / 1e6, 2), "unit": "MLUPS",
               "h2d_bytes_per_step": int(11 * cells * 8 / args.steps), "d2h_bytes_per_step": int(11 * cells * 8 / args.steps),
               "job": f"ek_run_from_host: upload 11 fields (pinned host) + init_equilibrium + {args.steps} steps + download 11 "
                      "fields in one call, copies overlapped with the first and last LBM pass",
               "seconds": round(dt, 4),
               "separate_calls": {"value": round(cells * args.steps / dt_plain / 1e6, 2), "seconds": round(dt_plain, 4)}}
    sim.close()

    # ---- the small configs (launch-latency bound): C1 as shipped and C2, step-only, state resident.
    # ek_step replays a CUDA graph of two coupled steps there.  The reference's figures for the same
    # configs are in the `extra` record of the --impl reference line.
    extra = {}
    if not args.no_extra:
        for name, steps2 in (("c1", 1000), ("c2", 400)):
            w2 = WORKLOADS[name]
            p2 = ek.default_params(NX=w2["NX"], NY=w2["NY"], NZ=w2["NZ"], **w2["over"])
            s2 = ek.Simulation(p2, device=local_rank)
            s2.init()
            s2.step(20)
            best = min(s2.step_timed(steps2) for _ in range(3))
            cells2 = w2["NX"] * w2["NY"] * w2["NZ"]
            extra[name] = {"workload": w2["name"], "grid": [w2["NX"], w2["NY"], w2["NZ"]], "steps": steps2,
                           "us_per_step": round(1e3 * best / steps2, 2), "mlups": round(cells2 * steps2 / best / 1e3, 1),
                           "graph_replays_per_call": int(s2.counter("graph_replays")) // 3,
                           "timing": "best of 3 x %d steps, CUDA events, macroscopic arrays written on the last step" % steps2}
            s2.close()

    cb = None if args.no_cpu_baseline else cpu_baseline()

    line = {"metric": "coupled_step_mlups", "value": round(mlups, 2), "unit": "MLUPS", "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": w["name"], "grid": [NX, NY, NZ], "stream_mode": args.stream_mode,
                       "zchunk": zchunk_used, "init": "reference start-up (501 PB iterations) %.2f s" % init_s,
                       "l2": "working set 14.5 GB of populations per step >> 126 MB L2 (no flush needed)",
                       "fields": "rho, u, c+, c-, T, E are written on the last step of an ek_step(n) call (what the "
                                 "reference's dumps read, main.cu:206); phi and c+ - c- every step",
                       "parallelism": "1 GPU"},
            "roofline": roofline, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "hbm_gbs_step": roofline["step_achieved"], "extra": extra}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
