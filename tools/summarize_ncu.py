"""Per-kernel table from `ncu -i X.ncu-rep --page raw --csv` output: duration, DRAM bytes, achieved GB/s against the
measured copy peak (MEASURED_PEAKS.json hbm_gbs), tensor-pipe share.  usage: summarize_ncu.py raw.csv [peak_gbs]"""
import csv
import json
import sys
from pathlib import Path

rows = list(csv.reader(open(sys.argv[1])))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    mp = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
    peak = json.loads(mp.read_text())["hbm_gbs"] if mp.exists() else 6540.0
h, units = rows[0], rows[1]
col = {c: i for i, c in enumerate(h)}


def val(r, name, default=0.0):
    i = col.get(name)
    if i is None or r[i] == "":
        return default
    x = float(r[i].replace(",", ""))
    u = units[i]
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
    return x * scale


print(f"{'kernel':58s} {'grid':>14s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>7s} {'%copy':>6s} {'dram%':>6s} {'tens%':>6s} {'regs':>5s}")
for r in rows[2:]:
    name = r[col["Kernel Name"]].replace("<unnamed>::", "").split("(")[0][:58]
    us = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / us / 1e3 if us else 0.0
    print(f"{name:58s} {r[col['Grid Size']]:>14s} {us:8.1f} {rd / 1e6:8.1f} {wr / 1e6:8.1f} {gbs:7.0f} {100 * gbs / peak:6.1f} "
          f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):6.1f} "
          f"{int(val(r, 'launch__registers_per_thread')):5d}")
