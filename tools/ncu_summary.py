"""Summarise ncu outputs into small text files for profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

  python tools/ncu_summary.py launches <launches.csv> [first_n]   -> per-kernel time shares of a launch list
  python tools/ncu_summary.py rep <file.ncu-rep>                  -> key metrics of each captured kernel
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(path, first_n=None, one_step=False):
    lines = open(path).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[i:]))
    if first_n:
        rows = rows[:first_n]
    if one_step:  # exactly one training step: from one embedding gather (the first kernel of a step) to the next
        marks = [k for k, r in enumerate(rows) if "embedding_kernel" in r["Kernel Name"]]
        if len(marks) >= 2:
            rows = rows[marks[-2]:marks[-1]]
            print(f"# one training step: launches {marks[-2]}..{marks[-1] - 1} of the capture (embedding gather to embedding gather)")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
        t = float(r["Metric Value"]) / 1e6
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    print(f"# {len(rows)} launches, {tot:.2f} ms summed kernel time (ncu: serialised, cold cache, unthrottled clocks - compare SHARES)")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{t:9.3f} ms {100 * t / tot:5.1f}%  x{c:4d}  avg {t / c * 1000:9.1f} us  {n[:110]}")


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        print("kernel:", d.get("Kernel Name", "?")[:140])
        for k in KEYS:
            if k in d:
                print(f"  {k} = {d[k]} {u[k]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
    elif sys.argv[1] == "step":
        launches(sys.argv[2], None, one_step=True)
    else:
        rep(sys.argv[2])
