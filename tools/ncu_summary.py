#!/usr/bin/env python
"""Compact per-launch summary (csv on stdout) of an ncu report captured with --set full.

usage: ncu_summary.py <report.ncu-rep>
"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
cols = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("gpu__time_duration.sum", "time_us"),
        ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("smsp__inst_executed.sum", "warp_instructions"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefront_pct"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct")]
idx = [(h.index(c), n) for c, n in cols if c in h]
units = rows[1]
w = csv.writer(sys.stdout)
w.writerow([n for _, n in idx])
for r in rows[2:]:
    vals = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0]
        elif n in ("dram_read_MB", "dram_write_MB"):
            scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(units[i], 1.0)
            v = "%.2f" % (float(v) * scale)
        elif n == "time_us":
            scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(units[i], 1.0)
            v = "%.2f" % (float(v) * scale)
        else:
            try:
                v = "%.1f" % float(v) if "." in v else v
            except ValueError:
                pass
        vals.append(v)
    w.writerow(vals)
