#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs,
and the hottest source lines by stall samples.  Usage: python tools_ncu_summary.py rep [n_lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__cluster_max_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_active.avg"]
for h, u, v in zip(hdr, units, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")) or h.startswith("sm__pipe_tensor") :
        print(f"{h:90s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if rows and rows[0] and rows[0][0] == "Kernel Name":
    rows = rows[1:]
if len(rows) > 2:
    h = rows[0]
    def col(name):
        for i, x in enumerate(h):
            if x == name: return i
        return None
    ci_src, ci_samp = col("Source"), col("# Samples") if col("# Samples") is not None else col("Samples")
    ci_addr = col("Address")
    print("columns:", h[:12], "...")
    if ci_samp is not None:
        data = []
        for r in rows[1:]:
            try: data.append((int(r[ci_samp]), r))
            except (ValueError, IndexError): pass
        tot = sum(d[0] for d in data) or 1
        data.sort(key=lambda x: -x[0])
        print(f"total samples {tot}")
        stall_cols = [(i, x) for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        for n, r in data[:nlines]:
            st = sorted(((int(r[i] or 0), x) for i, x in stall_cols), reverse=True)[:2]
            sts = " ".join(f"{x[6:]}={v}" for v, x in st if v)
            print(f"{n:7d} {100*n/tot:5.1f}%  {r[ci_addr][-5:]} {r[ci_src].strip()[:90]:90s} {sts}")
