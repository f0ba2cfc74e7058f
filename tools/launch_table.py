#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised times)."""
import collections
import csv
import io
import sys


def table(path, top=60):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [x for x in csv.DictReader(io.StringIO("".join(lines))) if x.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.OrderedDict()
    for x in rows:
        k = x["Kernel Name"].split("(")[0][:64]
        v = float(x["Metric Value"].replace(",", ""))
        if x.get("Metric Unit", "ns").startswith("us"):
            v *= 1e3
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(v for _, v in agg.values())
    out = ["%d launches, %.1f us in total" % (len(rows), tot / 1e3), "%-66s %6s %12s %7s" % ("kernel", "count", "us", "share")]
    for k, (c, v) in sorted(agg.items(), key=lambda t: -t[1][1])[:top]:
        out.append("%-66s %6d %12.1f %6.1f%%" % (k, c, v / 1e3, 100 * v / tot))
    lib = sum(c for k, (c, v) in agg.items() if "at::" in k or "cutlass" in k or "cublas" in k.lower() or "elementwise" in k)
    out.append("library (ATen / cuBLAS / cutlass) launches: %d" % lib)
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p)
        print(table(p))
