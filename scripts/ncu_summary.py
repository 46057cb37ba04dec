#!/usr/bin/env python
"""Turn an `ncu --set full` report into the small text evidence that is committed under profiles/.

    python scripts/ncu_summary.py REPORT.ncu-rep OUT_PREFIX [--units N --unit-name particle-step]
                                  [--workload cells4m --traffic-key dram_bytes_per_particle_step]

Writes OUT_PREFIX_summary.json (the metrics DESIGN.md / bench.py quote: duration, DRAM bytes,
instructions, issue / pipe utilisation, stall reasons per issue, registers, shared-memory
wavefronts) and OUT_PREFIX_raw.csv (the full `--page raw --csv` export of the first kernel, so
the binary report does not have to be tracked).  With --workload it also updates
profiles/traffic.json, which is where bench.py takes `roofline.traffic` from: DRAM bytes per
unit of work = (dram__bytes_read.sum + dram__bytes_write.sum) / units.
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "second": 1.0, "msecond": 1e-3,
         "usecond": 1e-6, "nsecond": 1e-9}

PICK = {
    "duration_s": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "warp_inst": "smsp__inst_executed.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "fmaheavy_pipe_pct": "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "xu_pipe_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "lsu_pipe_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "dyn_smem_bytes": "launch__shared_mem_per_block_dynamic",
    "shared_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "shared_wavefronts_ideal": "smsp__l1tex_data_pipe_lsu_wavefronts_mem_shared_ideal.sum",  # may be absent
}


def export_raw(report: str) -> str:
    return subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], check=True,
                          capture_output=True, text=True).stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out_prefix")
    ap.add_argument("--units", type=float, default=None, help="units of work in the captured launch")
    ap.add_argument("--unit-name", default="particle-step")
    ap.add_argument("--workload", default=None, help="bench.py workload this capture belongs to")
    ap.add_argument("--traffic-key", default="dram_bytes_per_particle_step")
    ap.add_argument("--note", default="")
    args = ap.parse_args()

    raw = export_raw(args.report)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(args.out_prefix + "_raw.csv", "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit", "value"])
        for h, u, v in zip(hdr, units, vals):
            w.writerow([h, u, v])

    def get(name):
        if name not in hdr:
            return None
        i = hdr.index(name)
        try:
            x = float(vals[i].replace(",", ""))
        except ValueError:
            return vals[i]
        return x * SCALE.get(units[i], 1.0)

    out = {"report": os.path.basename(args.report), "kernel": get("Kernel Name"), "note": args.note}
    for k, m in PICK.items():
        v = get(m)
        if v is not None:
            out[k] = v
    stalls = {}
    for h in hdr:
        pre, suf = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"
        if h.startswith(pre) and h.endswith(suf):
            v = get(h)
            if isinstance(v, float) and v >= 0.05:
                stalls[h[len(pre):-len(suf)]] = round(v, 3)
    out["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    if out.get("dram_read_bytes") is not None:
        out["dram_bytes"] = out["dram_read_bytes"] + out["dram_write_bytes"]
    if args.units:
        out["units"] = args.units
        out["unit_name"] = args.unit_name
        out["dram_bytes_per_unit"] = out["dram_bytes"] / args.units
        out["warp_inst_per_unit"] = out["warp_inst"] / args.units
        out["ns_per_unit"] = 1e9 * out["duration_s"] / args.units
    with open(args.out_prefix + "_summary.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))

    if args.workload and args.units:
        path = os.path.join(ROOT, "profiles", "traffic.json")
        try:
            with open(path) as f:
                tr = json.load(f)
        except Exception:
            tr = {}
        tr[args.workload] = {args.traffic_key: out["dram_bytes_per_unit"],
                             "source": f"{os.path.basename(args.out_prefix)}_summary.json: dram__bytes_read.sum + "
                                       f"dram__bytes_write.sum of one ncu --set full launch / {args.units:.0f} "
                                       f"{args.unit_name}s (report {os.path.basename(args.report)})"}
        with open(path, "w") as f:
            json.dump(tr, f, indent=1)


if __name__ == "__main__":
    sys.exit(main())
