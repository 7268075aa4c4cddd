#!/usr/bin/env python
"""Selected metrics of the raw page of an ncu report as JSON: `python tools/ncu_raw_summary.py report.ncu-rep out.json "description"`."""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
desc = sys.argv[3] if len(sys.argv) > 3 else ""
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__inst_executed_op_tma_ld.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "dram__sectors_read.sum", "dram__sectors_write.sum"]
res = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "", "what": desc}
for w in want:
    if w in hdr:
        i = hdr.index(w)
        res[w] = {"value": vals[i], "unit": units[i]}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res)[:400])
