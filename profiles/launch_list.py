"""Summarise `ncu --metrics gpu__time_duration.sum --csv` output: launches, total and average device time per kernel."""
import csv
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iu], 1e-6)
        name = r[ik].split("(")[0]
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + ms)
    total = sum(t for _, t in agg.values())
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:80]:80s} n={n:4d} total={t:9.2f} ms  avg={t / n:8.3f} ms  share={100 * t / total:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
