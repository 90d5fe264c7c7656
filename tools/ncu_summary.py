#!/usr/bin/env python
"""Turn the ncu exports of tools/gpu_job_profile.sh (gpurun_out/<tag>_*) into the files kept under profiles/:

  python tools/ncu_summary.py r02z [--sync 4]

  profiles/<tag>_launches.csv                       launch list (gpu__time_duration per launch) of a short bench.py run
  profiles/<tag>_ncu_<kernel>_{6,4}sweeps_details.txt   `--set full` details pages of the dominant kernel
  profiles/<tag>_ncu_k_spmm_dc_details.txt
  profiles/<tag>_ncu_key_metrics.json               the numbers quoted in DESIGN.md / profiles/*_notes.md
  profiles/dominant_kernel_traffic.json             DRAM bytes per launch of the dominant kernel = fixed + per_sweep x sweeps,
                                                    read by bench.py for roofline.traffic (refused there when the sweep kernel
                                                    or the colour count differ from the ones profiled)
"""
from __future__ import annotations

import argparse
import csv
import json
import shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "lts__t_requests_srcunit_tex_op_read.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw_rows(path: Path):
    rows = list(csv.reader(path.open()))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        ent = {"Kernel Name": d.get("Kernel Name")}
        for k in KEYS:
            if k in d:
                ent[k] = d[k] + (" " + u[k] if u.get(k) else "")
        out.append(ent)
    return out


def to_bytes(text: str) -> float:
    val, unit = text.split()
    return float(val) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--sync", type=int, default=4)
    ap.add_argument("--kernel", default="k_gs_tma")
    args = ap.parse_args()
    tag = args.tag
    shutil.copy(OUT / f"{tag}_launches.csv", PROF / f"{tag}_launches.csv")
    for m, sw in ((7, 6), (5, 4)):
        shutil.copy(OUT / f"{tag}_gs_m{m}_details.txt", PROF / f"{tag}_ncu_{args.kernel}_{sw}sweeps_details.txt")
    shutil.copy(OUT / f"{tag}_spmm_details.txt", PROF / f"{tag}_ncu_k_spmm_dc_details.txt")
    key = {f"{args.kernel}_6_sweeps": raw_rows(OUT / f"{tag}_gs_m7_raw.csv"), f"{args.kernel}_4_sweeps": raw_rows(OUT / f"{tag}_gs_m5_raw.csv"),
           "k_spmm": raw_rows(OUT / f"{tag}_spmm_raw.csv")}
    (PROF / f"{tag}_ncu_key_metrics.json").write_text(json.dumps(key, indent=1))
    b6 = sum(to_bytes(key[f"{args.kernel}_6_sweeps"][0][k]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    b4 = sum(to_bytes(key[f"{args.kernel}_4_sweeps"][0][k]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    plain = json.loads((OUT / f"{tag}_plain.json").read_text().strip().splitlines()[-1])
    per = (b6 - b4) / 2.0
    traffic = {"workload": "1m16", "family": "precond", "kernels": {f"sync{args.sync}": {
        "kernel": key[f"{args.kernel}_6_sweeps"][0]["Kernel Name"], "colours": plain["config"]["precond_colors"],
        "dram_bytes_per_sweep": per, "dram_bytes_fixed": b6 - 6 * per, "measured": {"4_sweeps": b4, "6_sweeps": b6},
        "source": f"ncu --set full --clock-control none (tools/gpu_job_profile.sh {tag}, profiles/{tag}_ncu_{args.kernel}_*_details.txt, "
                  f"{tag}_ncu_key_metrics.json): dram__bytes_read.sum + dram__bytes_write.sum of one launch with 6 sweeps per cycle "
                  "(precond_steps=7) and with 4 (precond_steps=5); linear in the sweep count, bench.py reports fixed + per_sweep x "
                  "(average sweeps per launch of the device-planned cycles)"}}}
    (PROF / "dominant_kernel_traffic.json").write_text(json.dumps(traffic, indent=1))
    print(json.dumps(traffic["kernels"], indent=1))


if __name__ == "__main__":
    main()
