#!/usr/bin/env python
"""Generates tests/golden/ref_snapshot_graphs.json from the reference's own insta snapshots
(/root/reference/src/tests/snapshots/*.snap, produced by the reference Writer under src/tests/writer.rs).

Run in the build container only (it reads /root/reference, which does not exist on the GPU box); the JSON it
writes is committed.  What is kept: the metadata line (dimensions, distance, entry points, max_level, item ids) and
every `Links` node in key order.  Keys sort by (item, layer) (big-endian KeyCodec, src/key.rs:54-66), so the k-th
`Links i` line of an item is its layer k-1.  Item vectors are printed truncated (4 decimals, first 10 dimensions)
in the snapshots and are NOT usable as golden vectors; the tests attach their own seeded vectors to these
reference-built topologies.
"""
import glob
import json
import os
import re

SNAP_DIR = "/root/reference/src/tests/snapshots"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_snapshot_graphs.json")


def parse(path):
    meta = None
    links = {}
    for line in open(path):
        line = line.strip()
        m = re.match(r"Root: Metadata \{ dimensions: (\d+), items: RoaringBitmap<(.*?)>, distance: \"(.*?)\", entry_points: \[(.*?)\], max_level: (\d+) \}", line)
        if m:
            meta = dict(dimensions=int(m.group(1)), items=m.group(2), distance=m.group(3),
                        entry_points=[int(x) for x in m.group(4).split(",") if x.strip()], max_level=int(m.group(5)))
            continue
        m = re.match(r"Links (\d+): Links\(Links \{ links: RoaringBitmap<\[(.*?)\]> \}\)", line)
        if m:
            item = int(m.group(1))
            nb = [int(x) for x in m.group(2).split(",") if x.strip()]
            links.setdefault(item, []).append(nb)   # position in the list == layer
    m = re.match(r"(\d+) values between (\d+) and (\d+)", meta["items"])
    if m:
        n, lo, hi = int(m.group(1)), int(m.group(2)), int(m.group(3))
        assert hi - lo + 1 == n, "sparse item set printed in summary form"
        items = list(range(lo, hi + 1))
    else:
        items = [int(x) for x in meta["items"].strip("[]").split(",") if x.strip()]
    meta["items"] = items
    return dict(source=os.path.relpath(path, "/root/reference"), metadata=meta,
                links=[[item, layer, nb] for item in sorted(links) for layer, nb in enumerate(links[item])])


def main():
    graphs = [parse(p) for p in sorted(glob.glob(os.path.join(SNAP_DIR, "*.snap")))]
    json.dump(dict(generated_by="tests/golden/make_reference_fixtures.py", graphs=graphs), open(OUT, "w"), separators=(",", ":"))
    for g in graphs:
        print(g["source"], len(g["metadata"]["items"]), "items,", len(g["links"]), "links nodes, max_level", g["metadata"]["max_level"])


if __name__ == "__main__":
    main()
