"""Markdown tables from ncu exports (run here, on the .ncu-rep / .csv files gpurun brought back).

  python tools/summarize_ncu.py launches gpurun_out/launches.csv         # per-kernel time shares of one step
  python tools/summarize_ncu.py full gpurun_out/prof_x.ncu-rep            # key counters of every captured launch
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

FULL_METRICS = OrderedDict([
    ("gpu__time_duration.sum", "time us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
])


def short(name):
    name = name.replace("void ", "")
    return name.split("(")[0][:48]


def launches(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= mv:
            continue
        t = float(r[mv].replace(",", "")) / 1000.0        # ns -> us
        a = agg.setdefault(short(r[kn]), [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if t / total >= 0.004:
            print("| `%s` | %d | %.1f | %.3f |" % (k, n, t, t / total))
    print("\n%d launches, %.1f us in total (cold-cache, serialised: shares only)" % (sum(v[0] for v in agg.values()), total))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), label) for m, label in FULL_METRICS.items() if m in hdr]
    kn = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join("%s%s" % (lab, (" [" + units[i] + "]") if units[i] not in ("", "%") else "")
                                     for i, lab in cols) + " |")
    print("|---" * (len(cols) + 1) + "|")
    for r in rows[2:]:
        vals = []
        for i, _ in cols:
            try:
                vals.append("%.4g" % float(r[i].replace(",", "")))
            except ValueError:
                vals.append(r[i])
        print("| `%s` | " % short(r[kn]) + " | ".join(vals) + " |")


def traffic(path, mode, label, rows, index="0"):
    """dram__bytes_read.sum + dram__bytes_write.sum of launch number `index` in the capture -> profiles/traffic.json,
    which bench.py reads for roofline.traffic (per row x rows):
        python tools/summarize_ncu.py traffic gpurun_out/x.ncu-rep bf16 'conv_wgrad[0]' 2048 [launch index]"""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rws = list(csv.reader(io.StringIO(out)))
    hdr, units = rws[0], rws[1]
    r = rws[2 + int(index)]
    def val(name):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    total = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = os.path.join(root, "profiles", "traffic.json")
    d = json.load(open(dst)) if os.path.exists(dst) else {}
    d.setdefault(mode, {})[label] = {"dram_bytes": total, "rows": int(rows), "kernel": short(r[hdr.index("Kernel Name")]),
                                     "source": "ncu --set full, %s (launch %s: %s rows)" % (os.path.basename(path), index, rows)}
    json.dump(d, open(dst, "w"), indent=1)
    print(label, mode, "%.1f MB at %s rows = %.1f KB per row" % (total / 1e6, rows, total / float(rows) / 1e3))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
