"""Stage timing of hrm::planners::HRM3D::plan() through the C++ host layer (HRM_HOST_TIMING=1 prints the per-stage wall clock).

    python tools/plan_timing.py [scene] [precision] [n] [Lx Ly] [cpu]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HRM_HOST_TIMING", "1")
from hrm_b200 import hostlib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "3D_superquadrics_cluttered_rabbit"
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10
sc = json.load(open(os.path.join(ROOT, "tests", "golden", "scenes.json")))[name]
if len(sys.argv) > 5:
    sc["Lx"], sc["Ly"] = int(sys.argv[4]), int(sys.argv[5])
rows = np.array(sc["arena"] + sc["obstacle"])
a = (len(sc["arena"]), len(sc["obstacle"]), rows, n, np.array(sc["robot"]), np.array(sc["quats"]), sc["lim"], sc["Lx"], sc["Ly"], 5, sc["end_points"][0],
     sc["end_points"][1])
for it in range(4):
    t0 = time.perf_counter()
    got = hostlib.plan3d(*a, per_orientation=True, precision=prec)
    wall = (time.perf_counter() - t0) * 1e3
    print("call %d: wall %.2f ms, plan() %.2f ms, solved=%s, V=%d, E=%d, edge tests=%d, launches=%d" % (
        it, wall, got["total_time_s"] * 1e3, got["solved"], len(got["vertex"]), len(got["edge"]), got["edge_tests"], got["gpu_launches"]), flush=True)
if len(sys.argv) > 6 and sys.argv[6] == "cpu":
    from oracle import hrmo

    t0 = time.perf_counter()
    ref = hrmo.plan3d(*a, per_orientation=True)
    print("cpu oracle (1 core): %.1f ms, solved=%s, V=%d, E=%d, identical=%s" % (
        (time.perf_counter() - t0) * 1e3, ref["solved"], len(ref["vertex"]), len(ref["edge"]),
        bool(np.array_equal(got["vertex"], ref["vertex"]) and np.array_equal(got["edge"], ref["edge"]))))
