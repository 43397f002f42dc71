"""Summarise ncu captures into profiles/ (tracked).  Usage:
    python tools/ncu_summary.py <tag> <full.ncu-rep> [<launches.csv>]
Writes profiles/<tag>_kernels.csv (one row per captured launch: time, DRAM bytes, pipe use, occupancy),
profiles/<tag>_launches.csv (per-kernel aggregate of the gpu__time_duration launch list) and updates
profiles/ncu_traffic.json (dram read+write bytes per launch, keyed by kernel, from the --set full capture)."""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep = sys.argv[1], sys.argv[2]
launches = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "-" else None
prefix = sys.argv[4] if len(sys.argv) > 4 else ""  # e.g. "bench:" for the capture of the headline launch
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
           "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           # pipe utilisation (which pipe is the wall of an issue-bound kernel) and the L1/L2 request traffic
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
           "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
cols = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name", "Block Size", "Grid Size") or "__" in h]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", f"{tag}_{'bench_' if prefix else ''}kernels.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in cols])
    for r in rows[2:]:
        w.writerow([r[i] for i in cols])

def to_bytes(v, unit):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

ki, ri, wi, gi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Grid Size")
traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
best = {}
for r in rows[2:]:
    name = r[ki].replace("void ", "").split("(")[0]
    b = to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
    if b > best.get(name, (0,))[0]:
        best[name] = (b, to_bytes(r[ri], units[ri]), to_bytes(r[wi], units[wi]))
for name, (b, rd, wr) in best.items():
    traffic[prefix + name] = {"dram_bytes": b, "read": rd, "write": wr, "capture": f"{tag}: {os.path.basename(rep)}, largest launch of this kernel"}
json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)

if launches:
    lr = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    h = lr[0]
    k, v = h.index("Kernel Name"), h.index("Metric Value")
    agg = OrderedDict()
    for r in lr[1:]:
        agg.setdefault(r[k], []).append(float(r[v].replace(",", "")))
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ns", "mean_ns", "min_ns", "max_ns", "share_of_listed_gpu_time"])
        tot = sum(sum(x) for x in agg.values())
        for name, x in agg.items():
            w.writerow([name, len(x), int(sum(x)), int(sum(x) / len(x)), int(min(x)), int(max(x)), round(sum(x) / tot, 4)])
print("wrote profiles for", tag)
