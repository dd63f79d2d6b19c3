#!/usr/bin/env python
"""Summarise an `ncu --csv --log-file` launch list (long format: one row per launch and metric) into one line per
kernel: launches, total / average device time, share of the listed time and, when collected, DRAM bytes per launch
and DRAM throughput.  usage: summarize_ncu.py launches.csv [title]"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    launches = OrderedDict()
    for r in csv.DictReader(lines):
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"], "block": r["Block Size"]})
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)  # -> us
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r["Metric Name"]] = v
    agg = defaultdict(lambda: defaultdict(float))
    for d in launches.values():
        name = d["name"].replace("void ", "").replace("unnamed>::", "")
        name = name.replace("(int)", "").replace("(bool)", "")
        name = re.sub(r"\[lambda.*", "[lambda]", name)
        name = re.sub(r"\((?:af::|const|double|long|int|unsigned|CUtensorMap|GemmArgs|SmallMlpArgs).*", "", name).strip()[:110]
        a = agg[name]
        a["n"] += 1
        for k, v in d.items():
            if isinstance(v, float):
                a[k] += v
        a["grid"] = d["grid"]
        a["block"] = d["block"]
    total = sum(a["gpu__time_duration.sum"] for a in agg.values()) or 1.0
    print(f"# {title}")
    print("# kernel | launches | total us | share of listed time | avg us | DRAM read MB/launch | DRAM write MB/launch | "
          "DRAM GB/s | dram throughput % of peak | regs | grid x block (last launch)")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n, t = a["n"], a["gpu__time_duration.sum"]
        rd, wr = a.get("dram__bytes_read.sum", 0) / n / 1e6, a.get("dram__bytes_write.sum", 0) / n / 1e6
        gbs = (rd + wr) * 1e6 / (t / n * 1e-6) / 1e9 if t else 0
        pct = a.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", 0) / n
        regs = a.get("launch__registers_per_thread", 0) / n
        print(f"{name} | {int(n)} | {t:.1f} | {t / total:.4f} | {t / n:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | {pct:.1f} | "
              f"{regs:.0f} | {a['grid']} x {a['block']}")


if __name__ == "__main__":
    main()
