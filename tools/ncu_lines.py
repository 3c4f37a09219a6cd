"""Warp-stall samples of one kernel aggregated per CUDA source line (ncu --import-source on, -lineinfo build).

usage: python tools/ncu_lines.py <rep> <kernel-regex> [launch-skip] [top-n]
"""
import collections
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
agg = collections.OrderedDict()
cur_file, hdr, ix = "?", None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    try:
        n = int(r[ix["# Samples"]] or 0)
    except ValueError:
        continue
    key = (cur_file, r[0])
    a = agg.setdefault(key, {"n": 0, "src": r[1], "stalls": collections.Counter(), "inst": 0})
    a["n"] += n
    try:
        a["inst"] += int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        pass
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                a["stalls"][h[6:]] += int(r[ix[h]] or 0)
            except ValueError:
                pass
tot = sum(a["n"] for a in agg.values())
print(f"# total samples {tot}, {len(agg)} source lines")
top = sorted(agg.items(), key=lambda kv: -kv[1]["n"])[:topn]
for (f, line), a in sorted(top, key=lambda kv: (kv[0][0], int(kv[0][1]) if kv[0][1].isdigit() else 0)):
    st = ", ".join(f"{k}:{v}" for k, v in a["stalls"].most_common(3))
    print(f"{f}:{line:>5s} {100.0 * a['n'] / max(tot, 1):5.1f}%  inst {a['inst']:>9d}  {a['src'].strip()[:100]:100s} {st}")
