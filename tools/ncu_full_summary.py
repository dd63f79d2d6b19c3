#!/usr/bin/env python
"""Summarise an `ncu --set full` report of the contraction launches of one bench step.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_full_summary.py raw.csv "title" [--traffic-json profiles/roofline_traffic.json]

One line per captured launch: device time, tensor-pipe activity, DRAM bytes, L2 hit rate, registers, shared-memory bank
conflicts and the top stall reasons; with --traffic-json also writes the file bench.py reads `roofline.traffic` from
(mean DRAM bytes per launch of the dual-accumulator contractions, the kernel `roofline` is quoted on).
"""
import csv
import json
import re
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    path, title = sys.argv[1], sys.argv[2]
    traffic_json = sys.argv[sys.argv.index("--traffic-json") + 1] if "--traffic-json" in sys.argv else None
    rows = list(csv.reader(open(path, newline="")))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        i = col[name]
        try:
            return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
        except ValueError:
            return float("nan")

    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print(f"# {title}")
    print("# kernel | grid | time ms | tensor pipe active % of elapsed | % of active | DRAM read MB | DRAM write MB | "
          "L2 hit % | regs | smem bank conflicts | top stall reasons (warps stalled per issue)")
    launches = []
    for r in data:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]].replace("void ", "").replace("(int)", "").replace("(bool)", "")).strip()
        stalls = sorted(((val(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                         for h in stall_cols), reverse=True)[:4]
        d = {"kernel": name, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
             "time_ms": val(r, "gpu__time_duration.sum"),
             "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
             "tensor_pipe_active_pct_of_elapsed": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
             "tensor_pipe_active_pct_of_active": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
             "registers": val(r, "launch__registers_per_thread"),
             "smem_bank_conflicts": val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
             "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
             "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")}
        launches.append(d)
        print(f"{name} | {d['grid']} | {d['time_ms']:.4f} | {d['tensor_pipe_active_pct_of_elapsed']:.1f} | "
              f"{d['tensor_pipe_active_pct_of_active']:.1f} | {d['dram_read_bytes'] / 1e6:.1f} | {d['dram_write_bytes'] / 1e6:.1f} | "
              f"{d['l2_hit_pct']:.1f} | {d['registers']:.0f} | {d['smem_bank_conflicts']:.0f} | "
              + ", ".join(f"{n} {v:.2f}" for v, n in stalls))
    # template arguments <ALAY, BLAY, NACC, CFG, MULTI>: NACC == 2 is the dual (value + R) contraction
    dual = [d for d in launches if re.search(r"gemm_dmma_kernel<\d+, \d+, 2,", d["kernel"])]
    if dual:
        mean = sum(d["dram_read_bytes"] + d["dram_write_bytes"] for d in dual) / len(dual)
        t = sum(d["time_ms"] for d in dual)
        w = sum(d["time_ms"] * d["tensor_pipe_active_pct_of_elapsed"] for d in dual) / t
        print(f"# dual contractions: {len(dual)} launches, {t:.3f} ms, mean DRAM bytes per launch {mean:.0f}, "
              f"time-weighted tensor pipe active {w:.1f} % of elapsed")
        if traffic_json:
            import os
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from autofunc_b200.build import kernel_source_sha
            json.dump({"source": title, "kernel_source_sha": kernel_source_sha(), "dram_bytes_per_launch": mean,
                       "launches": launches}, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main()
