"""Turns gpurun_out/<tag>_launches.csv + gpurun_out/<tag>_<kernel>.ncu-rep into the tracked summaries under profiles/:
    python tools/summarize_profile.py r01 r01_minksweep3d_f64 f64 16
(args: tag, ncu-rep stem, precision mode, layers per launch of the profiled command)"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep, prec, layers = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")

# ---- launch list -> per-kernel share -------------------------------------------------------------------------------
launch_csv = os.path.join(go, tag + "_launches.csv")
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 10]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1:]:
    tot[r[ki]] += float(r[vi].replace(",", ""))
    cnt[r[ki]] += 1
allns = sum(tot.values())
with open(os.path.join(out_dir, tag + "_launches.csv"), "w") as f:
    f.write(open(launch_csv).read())
lines = ["# %s launch list (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: compare SHARES)" % tag, "",
         "| kernel | launches | total ms | avg ms | share |", "|---|---:|---:|---:|---:|"]
for k in sorted(tot, key=lambda k: -tot[k]):
    lines.append("| `%s` | %d | %.3f | %.4f | %.1f%% |" % (k, cnt[k], tot[k] / 1e6, tot[k] / cnt[k] / 1e6, 100 * tot[k] / allns))

# ---- full capture of the dominant kernel ----------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", os.path.join(go, rep + ".ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u, v = rr[0], rr[1], rr[2]
m = {a: (b, c) for a, b, c in zip(h, u, v)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
lines += ["", "# %s: `ncu --set full` of the dominant kernel (%s, %d layers per launch)" % (rep, prec, layers), "", "| metric | unit | value |", "|---|---|---:|"]
for w in want:
    if w in m:
        lines.append("| %s | %s | %s |" % (w, m[w][0], m[w][1]))


def num(name):
    unit, val = m[name]
    x = float(val.replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}.get(unit, 1)


dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
S, M, L = 1001, 10000, 512
w = 4 if prec == "f32" else 8
alg = (3 * S * M * w + 2 * L * S * w) * layers
lines += ["", "DRAM traffic per launch: %.4g B (read %.4g + write %.4g) = %.4g B per layer; algorithmic bytes per layer %.4g => traffic / algorithmic = %.3f"
          % (dram, num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), dram / layers, alg / layers, dram / alg)]
open(os.path.join(out_dir, rep + "_summary.md"), "w").write("\n".join(lines) + "\n")
tj = os.path.join(out_dir, "traffic.json")
t = json.load(open(tj)) if os.path.exists(tj) else {}
t["%s_n100_S%d_L%d" % (prec, S, L)] = {"dram_bytes_per_layer": dram / layers, "source": rep + ".ncu-rep (ncu --set full, %d layers per launch)" % layers}
json.dump(t, open(tj, "w"), indent=1, sort_keys=True)
print("\n".join(lines))
