"""Top stalled SASS instructions of one kernel from an .ncu-rep source page.

usage: python tools/ncu_hot.py <rep> <kernel-regex> [launch-skip] [top-n]
"""
import csv
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[hi + 1:]:
    if r and r[0] == "Address":
        break  # a second view (e.g. high-level source) follows
    if len(r) == len(hdr):
        body.append(r)
tot = sum(int(r[ix["# Samples"]] or 0) for r in body)
print(f"# {rows[0][1][:80]}  total samples {tot}, {len(body)} SASS lines")
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]] or 0))[:topn]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for i in sorted(order):
    r = body[i]
    n = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{i:5d} {100.0*n/tot:5.1f}%  {r[ix['Source']][:90]:90s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
