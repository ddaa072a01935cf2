"""Summarise an `ncu --csv --metrics ...` launch list: per kernel, the mean of every metric.

    python tools/ncu_summary.py gpurun_out/launches.csv
"""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = csv.DictReader(lines)
    data = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        name = re.sub(r"ab::(rs::)?", "", name)[:48]
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        data.setdefault(name, collections.OrderedDict()).setdefault((r["Metric Name"], r["Metric Unit"]), []).append(v)
    for name, metrics in data.items():
        parts = []
        n = 0
        for (m, unit), vals in metrics.items():
            n = len(vals)
            mean = sum(vals) / len(vals)
            short = m.replace("gpu__time_duration.sum", "t").replace("dram__bytes_", "dram_").replace(".sum", "")
            short = short.replace("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%")
            if unit in ("ns", "nsecond"):
                parts.append("%s=%.1fus" % (short, mean / 1e3))
            elif unit == "byte":
                parts.append("%s=%.1fMB" % (short, mean / 1e6))
            elif unit == "Kbyte":
                parts.append("%s=%.1fMB" % (short, mean / 1e3))
            elif unit == "Mbyte":
                parts.append("%s=%.1fMB" % (short, mean))
            elif unit == "Gbyte":
                parts.append("%s=%.1fMB" % (short, mean * 1e3))
            else:
                parts.append("%s=%.2f%s" % (short, mean, unit))
        print("%-48s n=%-3d %s" % (name, n, "  ".join(parts)))


if __name__ == "__main__":
    main(sys.argv[1])
