"""Repeats HRM3D::plan() on the bundled cluttered map and prints the wall clock of every call (stall hunting aid):
    python tools/plan_repeat.py [calls] [scene] [Lx Ly]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hrm_b200 import hostlib  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 12
name = sys.argv[2] if len(sys.argv) > 2 else "3D_superquadrics_cluttered_rabbit"
sc = json.load(open(os.path.join(ROOT, "tests", "golden", "scenes.json")))[name]
if len(sys.argv) > 4:
    sc["Lx"], sc["Ly"] = int(sys.argv[3]), int(sys.argv[4])
rows = np.array(sc["arena"] + sc["obstacle"])
a = (len(sc["arena"]), len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), np.array(sc["quats"]), sc["lim"], sc["Lx"], sc["Ly"], 5, sc["end_points"][0],
     sc["end_points"][1])
out = []
for it in range(calls):
    t0 = time.perf_counter()
    got = hostlib.plan3d(*a, per_orientation=True, precision=0)
    out.append(((time.perf_counter() - t0) * 1e3, got["total_time_s"] * 1e3))
print("wall ms   :", " ".join("%.1f" % w for w, _ in out))
print("plan() ms :", " ".join("%.1f" % p for _, p in out))
