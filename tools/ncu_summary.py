#!/usr/bin/env python
"""Developer tool: per-launch summary of an .ncu-rep (`ncu -i rep --page raw --csv`): kernel, duration, DRAM bytes,
DRAM / tensor-pipe utilisation.  With --traffic-json writes profiles/roofline_traffic.json for bench.py from the
LARGEST conv_tc launch of the report."""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=float("nan")):
    try:
        return float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return default


def scale(name):
    u = units[col[name]] if name in col else ""
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e-3, "usecond": 1e-6, "second": 1.0,
            "nsecond": 1e-9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1.0)


out = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    t = get(r, "gpu__time_duration.sum") * scale("gpu__time_duration.sum")
    rd = get(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum")
    wr = get(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum")
    out.append({"kernel": r[col["Kernel Name"]][:70], "us": t * 1e6, "dram_read_MB": rd / 1e6, "dram_write_MB": wr / 1e6,
                "dram_GBps": (rd + wr) / t / 1e9 if t > 0 else float("nan"),
                "dram_pct": get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "tensor_pct": get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "warps_active_pct": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "regs": get(r, "launch__registers_per_thread")})
print(f"{'kernel':70s} {'us':>9s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>8s} {'dram%':>6s} {'tens%':>6s} {'occ%':>6s} {'regs':>5s}")
for o in out:
    print(f"{o['kernel']:70s} {o['us']:9.1f} {o['dram_read_MB']:9.1f} {o['dram_write_MB']:9.1f} {o['dram_GBps']:8.0f} "
          f"{o['dram_pct']:6.1f} {o['tensor_pct']:6.1f} {o['warps_active_pct']:6.1f} {o['regs']:5.0f}")
if "--traffic-json" in sys.argv:
    conv = [o for o in out if "conv_tc" in o["kernel"]]
    if conv:
        top = max(conv, key=lambda o: o["us"])
        path = sys.argv[sys.argv.index("--traffic-json") + 1]
        json.dump({"dram_bytes_per_launch": (top["dram_read_MB"] + top["dram_write_MB"]) * 1e6, "kernel": top["kernel"],
                   "launch_us_under_ncu": top["us"], "tensor_pipe_active_pct": top["tensor_pct"],
                   "source": f"ncu --set full capture {rep.split('/')[-1]} (dram__bytes_read.sum + dram__bytes_write.sum)"},
                  open(path, "w"), indent=1)
        print("wrote", path)
