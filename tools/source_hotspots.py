"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per source line:
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv ; python tools/source_hotspots.py src.csv [top]
Prints warp-instruction counts and stall samples (by reason) for the hottest lines of every file."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
cur_file, hdr = None, None
agg = defaultdict(lambda: defaultdict(float))
src_text = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or cur_file is None or len(r) != len(hdr):
        continue
    if r[0] == "":
        continue  # SASS rows belong to the last source line: ncu repeats line totals on the source row
    try:
        line = int(r[0])
    except ValueError:
        continue
    key = (cur_file, line)
    src_text[key] = r[1].strip()[:90]
    for i, h in enumerate(hdr):
        if h in ("Instructions Executed", "# Samples") or (h.startswith("stall_") and "Not Issued" not in h):
            try:
                agg[key][h] += float(r[i])
            except ValueError:
                pass
tot_inst = sum(v["Instructions Executed"] for v in agg.values())
tot_samp = sum(v["# Samples"] for v in agg.values())
print(f"total warp instructions {tot_inst:.0f}  samples {tot_samp:.0f}")
keys = sorted(agg, key=lambda k: -agg[k]["# Samples"])[:top]
for k in keys:
    v = agg[k]
    stalls = sorted(((v[h], h[6:]) for h in v if h.startswith("stall_") and v[h] > 0), reverse=True)[:3]
    st = " ".join(f"{n}:{100 * x / max(v['# Samples'], 1):.0f}%" for x, n in stalls)
    print(f"{k[0]}:{k[1]:4d} inst {100 * v['Instructions Executed'] / tot_inst:5.1f}%  samp {100 * v['# Samples'] / tot_samp:5.1f}%  [{st}]  {src_text[k]}")

# phase roll-up for hrm_b200/csrc/mink3d.cu (+ device helpers): edit the ranges when the kernel moves
if len(sys.argv) > 3 and sys.argv[3] == "phases":
    ranges = [("line-code helpers (phase A)", "mink3d.cu", 1, 149), ("entry + smem carve-up + reload", "mink3d.cu", 150, 230),
              ("prologue", "mink3d.cu", 231, 338), ("phase A", "mink3d.cu", 339, 548), ("face load/test lambdas (C, D)", "mink3d.cu", 549, 569),
              ("scan B1 (SWAR cell test)", "mink3d.cu", 570, 632), ("cell line sets / test_lines (C)", "mink3d.cu", 633, 654),
              ("scan loop + cell lists (B1)", "mink3d.cu", 655, 719), ("phase C (per-cell tests)", "mink3d.cu", 720, 751),
              ("heavy cells (C)", "mink3d.cu", 752, 785), ("phase D (re-test + intervals)", "mink3d.cu", 786, 831),
              ("rsqrt (phase A)", "devmath.cuh", 30, 50), ("triangle test (C, D)", "devmath.cuh", 51, 129),
              ("face_vertices / insert (C, D)", "devmath.cuh", 130, 200)]
    print("\nphase roll-up: share of warp instructions / of stall samples, top stall reasons")
    for name, f, a, b in ranges:
        inst = samp = 0.0
        st = defaultdict(float)
        for (ff, ln), v in agg.items():
            if ff == f and a <= ln <= b:
                inst += v["Instructions Executed"]
                samp += v["# Samples"]
                for h in v:
                    if h.startswith("stall_"):
                        st[h[6:]] += v[h]
        top3 = " ".join(f"{n}:{100 * x / max(samp, 1):.0f}%" for x, n in sorted(((x, n) for n, x in st.items()), reverse=True)[:4])
        print(f"  {name:30s} inst {100 * inst / tot_inst:5.1f}%  samples {100 * samp / tot_samp:5.1f}%  [{top3}]")
