#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the per-launch CSV kept under profiles/ and the per-kernel DRAM traffic that
bench.py reports as roofline.traffic.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_ncu_full.csv [profiles/traffic.json]
"""
import csv
import json
import subprocess
import sys

rep, out_csv = sys.argv[1], sys.argv[2]
out_json = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
idx = [hdr.index(c) if c in hdr else None for c in cols]
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[i] if i is not None else "" for i in idx])
    for r in data:
        w.writerow([r[i] if i is not None else "" for i in idx])


def to_bytes(v, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v.replace(",", "")) * scale


if out_json:
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    names = {"k_local_thin_in": "lift", "k_tc_local": "local", "k_tc_dft_fwd": "dft_fwd", "k_tc_mix_fwd": "mix", "k_tc_dft_inv": "dft_inv",
             "k_local_thin_out": "projection"}
    acc = {}
    for r in data:
        for k, short in names.items():
            if k in r[ik]:
                acc.setdefault(short, []).append(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]))
    json.dump({k: sum(v) / len(v) for k, v in acc.items()}, open(out_json, "w"), indent=1)
    print(open(out_json).read())
