#!/usr/bin/env python
"""Turn .ncu-rep files (brought back from the GPU box in gpurun_out/) into the small text/JSON summaries committed
under profiles/.  Needs only the `ncu` CLI (no GPU).

    python tools/ncu_summary.py traffic  <rep> <per-op-profile.json> <out.json>   # DRAM bytes per igemm launch, by op name
    python tools/ncu_summary.py details  <rep> <out.txt>                          # key metrics of every captured launch
    python tools/ncu_summary.py kernels  <rep> <out.json> [note]                  # mean time + DRAM bytes per kernel name
"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "launch__waves_per_multiprocessor",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    mode = sys.argv[1]
    if mode == "traffic":
        rep, prof, dst = sys.argv[2:5]
        hdr, units, rows = raw(rep)
        ops = [o for o in json.load(open(prof))["encode"] if o["kind"] == "igemm"]
        ib, it = hdr.index("dram__bytes.sum.per_second"), hdr.index("gpu__time_duration.sum")
        scale = {"Tbyte/s": 1e12, "Gbyte/s": 1e9, "Mbyte/s": 1e6, "Kbyte/s": 1e3, "byte/s": 1.0}[units[ib]]
        assert units[it] == "us" and len(rows) >= len(ops)
        table = {}
        for o, r in zip(ops, rows):
            us = float(r[it])
            table[o["name"]] = {"ncu_us": us, "dram_bytes": float(r[ib]) * scale * us * 1e-6, "kernel": r[hdr.index("Kernel Name")][:48],
                                "algorithmic_bytes": o["bytes"], "algorithmic_flops": o["flops"]}
        json.dump({"source": rep, "note": "ncu, one launch each, cold cache, --clock-control none; dram_bytes = "
                   "dram__bytes.sum.per_second x gpu__time_duration", "ops": table}, open(dst, "w"), indent=1)
        print("wrote", dst, len(table), "ops")
    elif mode == "kernels":
        rep, dst = sys.argv[2:4]
        note = sys.argv[4] if len(sys.argv) > 4 else ""
        hdr, units, rows = raw(rep)
        ib, it, ik = hdr.index("dram__bytes.sum.per_second"), hdr.index("gpu__time_duration.sum"), hdr.index("Kernel Name")
        scale = {"Tbyte/s": 1e12, "Gbyte/s": 1e9, "Mbyte/s": 1e6, "Kbyte/s": 1e3, "byte/s": 1.0}[units[ib]]
        assert units[it] == "us"
        acc = {}
        for r in rows:
            name = r[ik].split("(")[0].replace("fpnmt::", "").replace("void ", "").strip()
            a = acc.setdefault(name, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += float(r[it])
            a[2] += float(r[ib]) * scale * float(r[it]) * 1e-6
        json.dump({"source": "%s (ncu --set full, cold cache, --clock-control none) %s" % (rep, note),
                   "kernels": {k: {"launches": a[0], "mean_ncu_us": a[1] / a[0], "mean_dram_bytes": a[2] / a[0]}
                               for k, a in acc.items()}}, open(dst, "w"), indent=1)
        print("wrote", dst, {k: a[0] for k, a in acc.items()})
    elif mode == "details":
        rep, dst = sys.argv[2:4]
        hdr, units, rows = raw(rep)
        with open(dst, "w") as f:
            f.write("== %s\n" % rep)
            for r in rows:
                f.write(" kernel: %s grid %s block %s\n" % (r[hdr.index("Kernel Name")][:70], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
                for k in KEYS:
                    if k in hdr:
                        f.write("    %-90s %s %s\n" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        print("wrote", dst, len(rows), "launches")


if __name__ == "__main__":
    main()
