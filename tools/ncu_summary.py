"""Summarise an `ncu --set full` capture (raw CSV page) into the few numbers the roofline argument needs.
usage: ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python tools/ncu_summary.py X_raw.csv [kernel-key map json]"""
import csv
import json
import os
import sys


def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    names = {}
    gen = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "matx_b200", "csrc", "_gen")
    if os.path.isdir(gen):
        import re
        for fn in os.listdir(gen):
            if fn.endswith(".cu"):
                for key, sym in re.findall(r'\{"([^"]+)", \(const void \*\)(mxbk_[0-9a-f]+)\}', open(os.path.join(gen, fn)).read()):
                    names[sym] = key
    out = []
    for r in rows[2:]:
        g = lambda k: r[col[k]] if k in col else ""  # noqa: E731
        t_ns = f(g("gpu__time_duration.sum")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(units[col["gpu__time_duration.sum"]], 1)
        sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = f(g("dram__bytes_read.sum")) * sc.get(units[col["dram__bytes_read.sum"]], 1)
        wr = f(g("dram__bytes_write.sum")) * sc.get(units[col["dram__bytes_write.sum"]], 1)
        out.append({
            "kernel": names.get(g("Kernel Name"), g("Kernel Name")), "symbol": g("Kernel Name"), "grid": g("Grid Size"), "block": g("Block Size"),
            "time_us": round(t_ns / 1e3, 2), "dram_read_bytes": rd, "dram_write_bytes": wr,
            "dram_GBps": round((rd + wr) / t_ns, 1) if t_ns else None,
            "dram_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            "registers": g("launch__registers_per_thread"), "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "sm_throughput_pct": g("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            "inst_executed": g("smsp__inst_executed.sum"), "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
        })
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
