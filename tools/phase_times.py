"""Per-phase clock64() timing of the fused Minkowski+sweep kernel (profiling aid, not a bench):
    python tools/phase_times.py <f64|f64_strict|f32> <layers>"""
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
lib.hrmgpu_debug_phase_times.restype = C.c_longlong
lib.hrmgpu_debug_phase_times.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p]
rows = wl.scene_rows12()
ctx.set_scene3d(1, wl.N_OBS, wl.scene_rows17(rows), wl.N_GRID)
ctx.set_sweep(wl.LIM, wl.LX, wl.LY)
R = wl.body_rotations(wl.orientations(layers))
off = np.zeros((layers, 1, 3))
n_cta = layers * (wl.N_OBS + 1)
for rep in range(3):
    if rep == 2:
        lib.hrmgpu_debug_phase_times(ctx.h, 1, n_cta, None)
    ctx.build_layers(np.array(wl.BODY_SEMI), R, off, prec)
    print("rep", rep, "kernel ms", ctx.last_minksweep_ms())
out = np.zeros((n_cta, 8), dtype=np.int64)
lib.hrmgpu_debug_phase_times(ctx.h, 1, n_cta, out.ctypes.data_as(C.c_void_p))
if not os.environ.get("HRMGPU_CLASSIC"):
    # warp-specialised persistent kernel: one row per resident CTA, cycles of thread 0 of each role summed over its surfaces
    names = ["C: cell tests (+ heavy)", "P: wait (inputs + consumers)", "P: tables", "P: phase A", "C: wait for producers", "C: scan",
             "C: phase D", "surfaces"]
    rows_ = out[:min(n_cta, 2 * 148)]  # one row per resident CTA (2 per SM); rows 2g.. / 4g.. hold the per-warp scan details
    print("resident CTAs", len(rows_), "surfaces", rows_[:, 7].sum())
    for i, nm in enumerate(names):
        v = rows_[:, i] / (rows_[:, 7] if i < 7 else 1)
        print("   %-30s per surface: mean %10.1f  p50 %10.1f  max %10.1f" % (nm, v.mean(), np.percentile(v, 50), v.max()))
    g = len(rows_)
    ex = out[2 * g:3 * g] / rows_[:, 7:8]
    print("   scan per consumer warp (cycles per surface): SWAR tests %s | compaction %s | warp 0 waiting at the scan barrier %.0f"
          % (np.round(ex[:, 0:4].mean(axis=0)), np.round(ex[:, 4:8].mean(axis=0)), (out[4 * g:5 * g, 0] / rows_[:, 7]).mean()))
    e3 = out[3 * g:4 * g] / rows_[:, 7:8]
    print("   first SWAR section of a surface, per warp: %s | SWAR sections (chunks incl. re-entries) per surface: %s"
          % (np.round(e3[:, 0:4].mean(axis=0)), np.round(e3[:, 4:8].mean(axis=0), 2)))
    for blk in (0, 1, g - 1):
        print("   CTA %d hardware warp slots: producers %s consumers %s" % (blk, out[6 * g + blk].tolist(), out[7 * g + blk, :4].tolist()))
    ctx.close()
    sys.exit(0)
names = ["prologue", "phase A", "scan + tests", "of which C", "max B1 (warp)", "prologue: loads", "max D (thread)", "total"]
arena = (np.arange(n_cta) % (wl.N_OBS + 1)) == 0
for label, sel in (("obstacle CTAs", ~arena), ("arena CTAs", arena)):
    print(label, sel.sum())
    for i, nm in enumerate(names):
        v = out[sel, i]
        print("   %-13s mean %10.1f  p50 %10.1f  p99 %10.1f  max %10d" % (nm, v.mean(), np.percentile(v, 50), np.percentile(v, 99), v.max()))
ctx.close()
