"""Summarise an `ncu --csv --log-file` launch list (gpu__time_duration + dram bytes per launch) per kernel.

  python tools/ncu_summarize.py gpurun_out/launches.csv profiles/r01_ncu_step_summary   (-> .txt and .json)

The launch list is cold-cache and serialised (ncu replays every kernel), so the SHARES are what is comparable with
the CUDA-event numbers of bench.py, not the absolute times.
"""
import csv
import json
import re
import sys


def main(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        rows.append(r)
    agg = {}
    for r in rows:
        name = r.get("Kernel Name", "")
        metric = r.get("Metric Name", "")
        try:
            val = float(r.get("Metric Value", "0").replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "")
        short = re.sub(r"\(.*", "", name).replace("void ", "").strip()
        short = re.sub(r"<\(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(bool\)(\d)>", r"<\1,\2,\3,\4>", short)
        short = re.sub(r"<\(int\)(\d+), \(int\)(\d+)>", r"<\1,\2>", short)
        d = agg.setdefault(short, {"ids": set(), "time_ms": 0.0, "dram_read_MB": 0.0, "dram_write_MB": 0.0})
        d["ids"].add(r.get("ID"))
        if metric == "gpu__time_duration.sum":
            d["time_ms"] += val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        elif metric in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mb = val * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
            d["dram_read_MB" if "read" in metric else "dram_write_MB"] += mb
    total = sum(d["time_ms"] for d in agg.values()) or 1.0
    out = {}
    lines = []
    for k, d in sorted(agg.items(), key=lambda kv: -kv[1]["time_ms"]):
        n = len(d["ids"])
        out[k] = {"launches": n, "time_ms": round(d["time_ms"], 4), "share": round(d["time_ms"] / total, 4),
                  "dram_read_MB": round(d["dram_read_MB"], 2), "dram_write_MB": round(d["dram_write_MB"], 2),
                  "dram_MB_per_launch": round((d["dram_read_MB"] + d["dram_write_MB"]) / max(n, 1), 3)}
        lines.append(f"{d['time_ms']:9.3f} ms {100 * d['time_ms'] / total:5.1f}%  x{n:4d}  dram r/w "
                     f"{d['dram_read_MB']:9.1f}/{d['dram_write_MB']:9.1f} MB  {k[:90]}")
    with open(dst + ".txt", "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(dst + ".json", "w") as f:
        json.dump({"total_ms": round(total, 3), "kernels": out}, f, indent=1)
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
