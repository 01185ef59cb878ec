"""Condense `ncu -i report.ncu-rep --page raw --csv` (first kernel of the report) into the JSON summary kept under profiles/.
    python scripts/ncu_summary.py raw.csv "<kernel description>" "<source command>" "<workload>" algorithmic_bytes [extra_written_bytes]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, unit, val = rows[h], rows[h + 1], rows[h + 2]
m = dict(zip(hdr, val))
u = dict(zip(hdr, unit))


def f(name, scale=1.0):
    v = m.get(name)
    if v in (None, "", "n/a"):
        return None
    x = float(v.replace(",", ""))
    un = u.get(name, "")
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
            "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(un, 1.0)
    return x * mult * scale


rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
stalls = {}
for k in hdr:
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
        name = k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
        if name not in ("selected",):
            stalls[name] = round(float(m[k].replace(",", "")), 3)
top = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
alg = float(sys.argv[5])
out = {
    "kernel": sys.argv[2], "source": sys.argv[3], "workload": sys.argv[4],
    "kernel_name": m.get("Kernel Name"), "grid": m.get("Grid Size"), "block": m.get("Block Size"),
    "gpu__time_duration_ms": f("gpu__time_duration.sum"),
    "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": (rd or 0) + (wr or 0),
    "algorithmic_bytes_per_launch": alg, "traffic_over_algorithmic": round(((rd or 0) + (wr or 0)) / alg, 3),
    "registers_per_thread": f("launch__registers_per_thread"),
    "achieved_occupancy_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    "lsu_pipe_pct": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "l1tex_lsu_data_pipe_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "issue_slots_busy_pct": f("sm__inst_issued.avg.pct_of_peak_sustained_active") or f("smsp__issue_active.avg.pct"),
    "dram_throughput_pct": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    "inst_executed": f("smsp__inst_executed.sum"),
    "shared_bank_conflicts": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    "top_stalls_per_issue": top,
}
if len(sys.argv) > 6:
    out["note"] = sys.argv[6]
print(json.dumps(out, indent=1))
