"""Short, deterministic run of the hot path for profiling under ncu (never a bench number):
    python tools/prof_layers.py <f64|f64_strict|f32> <layers> <reps> [n_obs]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hrm_b200 import capi, workload as wl  # noqa: E402

prec = {"f64": capi.F64, "f64_strict": capi.F64_STRICT, "f32": capi.F32}[sys.argv[1]]
layers = int(sys.argv[2])
reps = int(sys.argv[3])
n_obs = int(sys.argv[4]) if len(sys.argv) > 4 else wl.N_OBS
rows = wl.scene_rows12(wl.N_OBS)[: 1 + n_obs]
ctx = capi.Context(0)
ctx.set_scene3d(1, n_obs, wl.scene_rows17(rows), wl.N_GRID)
ctx.set_sweep(wl.LIM, wl.LX, wl.LY)
R = wl.body_rotations(wl.orientations(layers * reps))
off = np.zeros((layers * reps, 1, 3))
for r in range(reps):
    ctx.build_layers(np.array(wl.BODY_SEMI), R[r * layers:(r + 1) * layers], off[r * layers:(r + 1) * layers], prec)
    print("rep", r, "minksweep ms", ctx.last_minksweep_ms(), "segments", ctx.dims()["segments"])
ctx.close()
