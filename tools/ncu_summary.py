#!/usr/bin/env python3
"""Summarise the ncu outputs of tools/profile_gpu.sh into profiles/ (run here, no GPU needed).

usage: ncu_summary.py TAG      reads gpurun_out/TAG_launches.csv and gpurun_out/TAG_full.ncu-rep
writes profiles/TAG_launches_by_kernel.csv, profiles/TAG_ncu_full_summary.json and refreshes
profiles/lbm_kernel_traffic.json (DRAM bytes per LBM launch, read by bench.py for roofline.traffic).
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "sm__cycles_elapsed.avg.per_second"]


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        n, t = agg.get(r[ik], (0, 0.0))
        agg[r[ik]] = (n + 1, t + v)
    out = os.path.join(ROOT, "profiles", f"{tag}_launches_by_kernel.csv")
    total = sum(t for _, t in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{n},{t:.1f},{t / total:.4f}\n')
    return out


def full(tag):
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_full.ncu-rep")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    res = []
    for r in rows[2:]:
        d = {"kernel": r[ik]}
        for m in KEEP:
            if m in hdr:
                j = hdr.index(m)
                d[m] = f"{r[j]} {units[j]}".strip()
        res.append(d)
    out = os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.json")
    with open(out, "w") as f:
        json.dump(res, f, indent=1)
    # DRAM bytes per lean LBM launch (mean over the captured even/odd launches)
    def gb(s):
        v, u = s.split()
        return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    lbm = [d for d in res if "ek_step_kernel" in d["kernel"]]
    # launches that also write the seven macroscopic arrays (FULL, second template argument) move 56 B/cell more
    # than B_alg counts: the roofline traffic is the plain launches' when the capture holds any
    plain = [d for d in lbm if re.search(r"ek_step_kernel<\d+, (0|false)", d["kernel"])]
    lbm = plain or lbm
    if lbm:
        rd = [gb(d["dram__bytes_read.sum"]) for d in lbm]
        wr = [gb(d["dram__bytes_write.sum"]) for d in lbm]
        tr = {"kernel": "ek_step_kernel (A-A even/odd, lean path)", "source": f"profiles/{tag}_ncu_full_summary.json "
              "(ncu --set full --clock-control none, 256^3)", "launches": len(lbm),
              "dram_bytes_read": round(sum(rd) / len(rd)), "dram_bytes_write": round(sum(wr) / len(wr)),
              "dram_bytes_per_launch": round((sum(rd) + sum(wr)) / len(lbm)),
              "algorithmic_bytes_per_launch": 256 ** 3 * 1744}
        with open(os.path.join(ROOT, "profiles", "lbm_kernel_traffic.json"), "w") as f:
            json.dump(tr, f, indent=1)
    return out


if __name__ == "__main__":
    tag = sys.argv[1]
    print(launches(tag))
    print(full(tag))
