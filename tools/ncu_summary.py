"""Summarise an `ncu --set full` report as a markdown table for profiles/ (one row per captured launch).

    python tools/ncu_summary.py gpurun_out/prof_main_r01b.ncu-rep [--title "..."] > profiles/r01_xxx.md

Reads the report through `ncu -i <rep> --page raw --csv` (no GPU needed).  Columns: duration, DRAM bytes
read / written (the `traffic` figure bench.py's roofline quotes), DRAM and SM throughput as % of peak,
achieved occupancy, registers / thread, shared memory / block, L2 hit rate, tensor-pipe utilisation.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram rd"),
    ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__shared_mem_per_block_static", "static smem"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor % (elapsed)"),
]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6,
              "usecond": 1, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}


def short(name: str) -> str:
    name = name.replace("void ", "").replace("<unnamed>::", "")
    p = name.find("(")
    return name[:p] if p > 0 else name


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    return rows[hdr], rows[hdr + 1], rows[hdr + 2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--title", default=None)
    ap.add_argument("--json", action="store_true", help="also print one JSON object per launch (for bench.py's traffic table)")
    a = ap.parse_args()
    names, units, data = load(a.rep)
    idx = {n: i for i, n in enumerate(names)}
    print("# %s\n" % (a.title or os.path.basename(a.rep)))
    print("Source: `%s` (`ncu --set full --clock-control none --import-source on`); one row per captured launch.\n" % os.path.basename(a.rep))
    present = [(m, t) for m, t in COLS if m in idx and any(r[idx[m]] not in ("", "n/a") for r in data)]
    print("| # | kernel | grid | block | " + " | ".join(t for _, t in present) + " | dram GB/s |")
    print("|---|---|---|---|" + "---|" * (len(present) + 1))
    js = []
    for r in data:
        if not r or len(r) < len(names):
            continue
        cells, rec = [], {"kernel": short(r[idx["Kernel Name"]]), "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        for m, t in present:
            v, u = r[idx[m]], units[idx[m]].split('/')[0]
            try:
                f = float(v.replace(",", ""))
            except ValueError:
                cells.append(v); continue
            if m == "gpu__time_duration.sum":
                f *= UNIT_SCALE.get(u, 1); rec["us"] = f; cells.append("%.1f us" % f)
            elif "bytes" in m or "shared_mem" in m:
                f *= UNIT_SCALE.get(u, 1); rec[t] = f
                cells.append("%.2f MB" % (f / 1e6) if f >= 1e5 else "%.0f B" % f)
            else:
                rec[t] = f; cells.append("%.1f" % f if "%" in t else "%.0f" % f)
        gbs = ""
        if "us" in rec and "dram rd" in rec and "dram wr" in rec and rec["us"] > 0:
            rec["dram_GBps"] = (rec["dram rd"] + rec["dram wr"]) / rec["us"] / 1e3
            gbs = "%.0f" % rec["dram_GBps"]
        print("| %s | `%s` | %s | %s | " % (r[idx["ID"]], rec["kernel"], rec["grid"], rec["block"]) + " | ".join(cells) + " | %s |" % gbs)
        js.append(rec)
    if a.json:
        print("\n```json")
        for rec in js:
            print(json.dumps(rec))
        print("```")


if __name__ == "__main__":
    main()
