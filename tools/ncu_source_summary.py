#!/usr/bin/env python
"""Condense the source page of an ncu report (`ncu --set full --import-source on`) into per-source-line shares of executed
instructions and stall samples: `python tools/ncu_source_summary.py report.ncu-rep [out.json] [top]`.
Reads `ncu -i report --page source --csv --print-source cuda,sass` (run where ncu is installed; no GPU needed)."""
import csv, io, json, os, subprocess, sys

rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file, hdr, kernel = None, None, None
lines = {}
stall_cols = {}
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "File Path":
        cur_file = os.path.basename(row[1]); continue
    if row[0] == "Function Name":
        kernel = row[1]; continue
    if row[0] == "Line No":
        hdr = row
        stall_cols = {i: h for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
        continue
    if hdr is None or row[0] == "" or not row[0].isdigit():
        continue   # SASS rows under a source line: already summed into the line's own row
    i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    try:
        inst, samp = int(row[i_inst]), int(row[i_samp])
    except ValueError:
        continue
    if inst == 0 and samp == 0:
        continue
    d = lines.setdefault(f"{cur_file}:{row[0]}", {"inst": 0, "samples": 0, "stalls": {}, "src": row[1].strip()[:110]})
    d["inst"] += inst; d["samples"] += samp
    for i, h in stall_cols.items():
        try:
            v = int(row[i])
        except ValueError:
            v = 0
        if v:
            d["stalls"][h] = d["stalls"].get(h, 0) + v
ti = sum(d["inst"] for d in lines.values()) or 1
ts = sum(d["samples"] for d in lines.values()) or 1
by_file, stall_tot = {}, {}
for k, d in lines.items():
    f = k.split(":")[0]
    a = by_file.setdefault(f, [0, 0]); a[0] += d["inst"]; a[1] += d["samples"]
    for h, v in d["stalls"].items():
        stall_tot[h] = stall_tot.get(h, 0) + v
res = {"kernel": kernel, "instructions_executed": ti, "stall_samples": ts,
       "by_file": {f: {"instructions_frac": round(a[0] / ti, 4), "stall_samples_frac": round(a[1] / ts, 4)} for f, a in sorted(by_file.items(), key=lambda kv: -kv[1][0])},
       "stall_reasons": {h: round(v / ts, 4) for h, v in sorted(stall_tot.items(), key=lambda kv: -kv[1]) if v / ts >= 0.005},
       "lines": {k: {"instructions_frac": round(d["inst"] / ti, 4), "stall_samples_frac": round(d["samples"] / ts, 4),
                     "top_stall": max(d["stalls"], key=d["stalls"].get) if d["stalls"] else None, "src": d["src"]}
                 for k, d in sorted(lines.items(), key=lambda kv: -(kv[1]["inst"] / ti + kv[1]["samples"] / ts))[:top]}}
if out:
    json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "lines"}, indent=1))
for k, v in res["lines"].items():
    print(f"{k:22s} inst {v['instructions_frac']:.4f} stall {v['stall_samples_frac']:.4f} {str(v['top_stall']):22s} {v['src'][:90]}")
