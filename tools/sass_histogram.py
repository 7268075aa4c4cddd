#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libhannoy_b200.so (cuobjdump -sass; runs without a GPU): the committed evidence that
the kernels are what DESIGN.md says — bulk async copies (UBLKCP / UTMALDG) completing on mbarriers (SYNCS.*), tcgen05 MMA
(UTC*MMA) with TMEM loads (LDTM), global atomics (ATOMG / REDG), and no legacy tensor path (HMMA).

  python tools/sass_histogram.py [out.json]
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "hannoy_b200", "libhannoy_b200.so")
KEY = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "HMMA", "ATOMG", "ATOMS", "REDG", "RED",
       "ATOM", "LDG", "STG", "LDS", "STS", "LDGSTS", "SHFL", "POPC", "FFMA", "FADD", "FMUL", "BAR", "MEMBAR", "CCTL", "VOTE", "MATCH", "REDUX", "LDL", "STL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            name = re.sub(r"\(.*", "", name).replace("hb::(anonymous namespace)::", "hb::").strip()
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            op = m.group(1)
            kernels[name]["total"] += 1
            kernels[name][op.split(".")[0]] += 1
            if op.split(".")[0] in ("SYNCS", "UBLKCP", "UTMALDG", "LDTM", "UTCHMMA"):
                kernels[name][op] += 1
    res = {}
    for k, c in kernels.items():
        row = {"instructions": c["total"]}
        for op in sorted(c):
            base = op.split(".")[0]
            if op != "total" and (base in KEY or op in KEY) and c[op]:
                row[op] = c[op]
        res[k] = row
    text = json.dumps(res, indent=1)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
