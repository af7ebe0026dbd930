"""Summarise an `ncu --set full` capture of decode_step_kernel into profiles/: python tools/ncu_step_summary.py rep workload tag

Writes profiles/<tag>_step_kernel_raw.csv (selected raw metrics) and updates profiles/step_kernel_ncu.json (DRAM bytes per
launch, read by bench.py for roofline.traffic)."""
import csv
import json
import os
import subprocess
import sys

rep, workload, tag = sys.argv[1:4]
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
sel = [(h, units[i], vals[i]) for i, h in enumerate(hdr) if h in want]
out = os.path.join(REPO, "profiles", f"{tag}_step_kernel_raw.csv")
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    w.writerows(sel)
d = {h: (u, v) for h, u, v in sel}


def to_bytes(name):
    u, v = d[name]
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


path = os.path.join(REPO, "profiles", "step_kernel_ncu.json")
js = json.load(open(path)) if os.path.exists(path) else {}
js[workload] = {"dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
                "gpu_time_us_under_ncu": float(d["gpu__time_duration.sum"][1]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[d["gpu__time_duration.sum"][0]], "source": os.path.basename(rep), "tag": tag}
json.dump(js, open(path, "w"), indent=1, sort_keys=True)
print(out, js[workload])
