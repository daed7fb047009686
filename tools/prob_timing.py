"""Stage timing of hrm::planners::ProbHRM3D::plan() on the home map with the snake robot (HRM_HOST_TIMING=1 prints the stages):
    python tools/prob_timing.py [calls] [batch]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HRM_HOST_TIMING", "1")
from hrm_b200 import hostlib  # noqa: E402
from tests.test_oracle_kat import GOLDEN, prob_samples  # noqa: E402

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sc = json.load(open(os.path.join(ROOT, "tests", "golden", "scenes.json")))["3D_superquadrics_home_snake"]
rows = np.array(sc["arena"] + sc["obstacle"])
a = (len(sc["arena"]), len(sc["obstacle"]), rows, 10, np.array(sc["robot"]), os.path.join(GOLDEN, "snake.urdf"), prob_samples(30, 3), sc["lim"],
     sc["Lx"], sc["Ly"], 5, sc["end_points"][0], sc["end_points"][1])
for it in range(calls):
    t0 = time.perf_counter()
    got = hostlib.plan_prob3d(*a, per_orientation=True, precision=0, batch=batch)
    print("call %d: wall %.1f ms, plan() %.1f ms, slices %d, V=%d, E=%d, solved=%s, launches=%d" % (
        it, (time.perf_counter() - t0) * 1e3, got["total_time_s"] * 1e3, got["slices"], len(got["vertex"]), len(got["edge"]), got["solved"],
        got["gpu_launches"]), flush=True)
