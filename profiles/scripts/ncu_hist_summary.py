#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of the histogram kernel launches of ONE boosting round into hist_traffic.json
(the file bench.py reads roofline.traffic from).  usage: ncu_hist_summary.py <report.ncu-rep> <out.json> <rows> <cols>"""
import csv
import io
import json
import subprocess
import sys

rep, out, rows, cols = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rd = list(csv.reader(io.StringIO(raw)))
hdr = next(i for i, r in enumerate(rd) if r and r[0] == "ID")
H, units = rd[hdr], rd[hdr + 1]


def col(name):
    return H.index(name) if name in H else None


def num(r, name, scale=1.0):
    i = col(name)
    if i is None or r[i] in ("", "n/a"):
        return None
    v = float(r[i].replace(",", ""))
    u = units[i]
    if name.endswith("bytes_read.sum") or name.endswith("bytes_write.sum"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if name == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    return v * scale


L = []
for r in rd[hdr + 2:]:
    if len(r) < len(H):
        continue
    L.append({
        "kernel": r[col("Kernel Name")][:60],
        "duration_us": num(r, "gpu__time_duration.sum"),
        "dram_read_MB": (num(r, "dram__bytes_read.sum") or 0) / 1e6,
        "dram_write_MB": (num(r, "dram__bytes_write.sum") or 0) / 1e6,
        "dram_pct_peak": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "l1tex_data_pipe_pct": num(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "smem_atom_wavefronts": num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum")
        or num(r, "smsp__inst_executed_op_shared_atom.sum"),
        "issue_active_pct": num(r, "sm__inst_issued.avg.pct_of_peak_sustained_active")
        or num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "regs": num(r, "launch__registers_per_thread"),
        "grid": num(r, "launch__grid_size"),
    })
tot = sum((x["dram_read_MB"] + x["dram_write_MB"]) * 1e6 for x in L)
json.dump({"source": "ncu --set full --clock-control none --import-source on -k regex:hist_build of one boosting round "
                     "(profiles/scripts/r02_batch3.sh); %d x %d" % (rows, cols),
           "launches": L, "dram_bytes_per_launch": tot / max(1, len(L)),
           "metric_columns_present": [h for h in H if "wavefront" in h or "atom" in h][:40]}, open(out, "w"), indent=1)
print("launches", len(L), "dram bytes/launch", tot / max(1, len(L)))
