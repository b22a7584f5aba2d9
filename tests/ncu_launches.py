"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and mean duration, share per kernel."""
import csv
import re
import sys


def main(path, skip_first=0, count=None):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, order = {}, []
    for r in rows[1 + skip_first:][:count]:
        if len(r) <= vi:
            continue
        name = re.sub(r"<unnamed>::|\(anonymous namespace\)::", "", r[ki])
        name = re.sub(r"^void ", "", name).split("(")[0]
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)     # -> us
        if name not in agg:
            agg[name] = [0, 0.0]
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += v
    total = sum(a[1] for a in agg.values()) or 1.0
    print("%-64s %9s %12s %10s %7s" % ("kernel", "launches", "total us", "avg us", "share"))
    for name in order:
        n, t = agg[name]
        print("%-64s %9d %12.1f %10.1f %6.1f%%" % (name[:64], n, t, t / n, 100 * t / total))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else None)
