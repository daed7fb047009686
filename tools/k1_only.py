"""K1 alone (boundary points, no sweep) next to the fused kernel on the bench workload (profiling aid, not a bench):
    python tools/k1_only.py <f64|f32> <layers>"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hrm_b200 import capi, workload as wl  # noqa: E402

prec = {"f64": capi.F64, "f64_strict": capi.F64_STRICT, "f32": capi.F32}[sys.argv[1]]
layers = int(sys.argv[2])
ctx = capi.Context(0)
lib = ctx.lib
lib.hrmgpu_debug_k1_only_ms.restype = C.c_double
lib.hrmgpu_debug_k1_only_ms.argtypes = [C.c_void_p]
rows = wl.scene_rows12()
ctx.set_scene3d(1, wl.N_OBS, wl.scene_rows17(rows), wl.N_GRID)
ctx.set_sweep(wl.LIM, wl.LX, wl.LY)
R = wl.body_rotations(wl.orientations(layers))
off = np.zeros((layers, 1, 3))
for rep in range(4):
    ctx.build_layers(np.array(wl.BODY_SEMI), R, off, prec)
    fused = ctx.last_minksweep_ms()
    k1 = lib.hrmgpu_debug_k1_only_ms(ctx.h)
    w = 4 if prec == capi.F32 else 8
    gb = layers * (wl.N_OBS + 1) * 3 * wl.N_GRID ** 2 * w / 1e9
    print("rep %d fused %.3f ms  K1 only %.3f ms (%.0f GB/s of mesh stores)" % (rep, fused, k1, gb / (k1 * 1e-3)))
ctx.close()
