#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the small JSON kept under profiles/: per kernel launch the duration, DRAM
bytes, pipe utilisation, issue / occupancy and shared-memory counters the design notes quote.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.json "free-text source line" [units_per_launch unit_name]"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def main():
    rep, out, source = sys.argv[1], sys.argv[2], sys.argv[3]
    units = float(sys.argv[4]) if len(sys.argv) > 4 else None
    unit_name = sys.argv[5] if len(sys.argv) > 5 else "unit"
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head, unit_row = rows[0], rows[1]
    tensor_like = [h for h in head if h.startswith("sm__pipe_tensor") and h.endswith("pct_of_peak_sustained_active") and h not in KEYS]
    kn = head.index("Kernel Name")
    kernels = []
    for r in rows[2:]:
        k = {"kernel": r[kn].split("(")[0]}
        for key in KEYS + tensor_like:
            if key in head:
                i = head.index(key)
                try:
                    k[key] = {"value": float(r[i].replace(",", "")), "unit": unit_row[i]}
                except ValueError:
                    pass
        if units and "dram__bytes_read.sum" in k:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            b = sum(k[m]["value"] * scale.get(k[m]["unit"], 1.0) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum") if m in k)
            k["dram_bytes_per_" + unit_name] = b / units
        kernels.append(k)
    json.dump({"source": source, "kernels": kernels}, open(out, "w"), indent=1)
    for k in kernels:
        keep = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
        print(k["kernel"][:50], {m.split(".")[0]: round(k[m]["value"], 2) for m in keep if m in k}, {m: round(v, 2) for m, v in k.items() if m.startswith("dram_bytes_per")})


main()
