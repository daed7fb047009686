"""Smallest case that shows the fused kernel at the bench shape (for ncu): `layers` C-layers of the bench scene, built 3 times.
    python tools/ncu_case.py <f64|f64_strict|f32> <layers> [scaling|freespace]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hrm_b200 import capi, workload as wl  # noqa: E402

prec = {"f64": capi.F64, "f64_strict": capi.F64_STRICT, "f32": capi.F32}[sys.argv[1]]
layers = int(sys.argv[2])
name = sys.argv[3] if len(sys.argv) > 3 else "scaling"
ctx = capi.Context(0)
rows = wl.scene_rows12(workload=name)
ctx.set_scene3d(1, wl.N_OBS, wl.scene_rows17(rows), wl.N_GRID)
ctx.set_sweep(wl.LIM, wl.LX, wl.LY)
R = wl.body_rotations(wl.orientations(layers))
off = np.zeros((layers, 1, 3))
reps = int(os.environ.get("HRM_CASE_REPS", "3"))
ms = []
for rep in range(reps):
    ctx.build_layers(np.array(wl.BODY_SEMI), R, off, prec)
    ctx.synchronize()
    ms.append(ctx.last_minksweep_ms())
    print("rep", rep, "fused kernel ms", ms[-1], "segments ms", ctx.last_segments_ms(), "segments", ctx.dims()["segments"])
if reps > 3:
    print("steady state (reps 2..): min %.3f median %.3f ms" % (min(ms[2:]), float(np.median(ms[2:]))))
ctx.close()
