"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python tools/summarize_launches.py gpurun_out/launches.csv [steps] > profiles/rNN_launches.txt"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", r[ki]))[:90]
        ns = float(r[vi].replace(",", ""))
        agg[name][0] += 1
        agg[name][1] += ns
        total += ns
    n = len(rows) - 1
    print("%d launches over %d step(s): %.2f ms of kernel time per step (ncu: cold-cache, serialised; compare shares)"
          % (n, steps, total / 1e6 / steps))
    print("%10s %7s %7s  %s" % ("us/step", "n/step", "share", "kernel"))
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%10.1f %7.1f %6.2f%%  %s" % (t / 1e3 / steps, c / steps, 100 * t / total, k))


if __name__ == "__main__":
    main()
