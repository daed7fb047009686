"""Tiny deterministic run of every precision mode and both fused kernels (for compute-sanitizer): bundled cluttered map,
n = 10 / 23 (one-surface kernel) and n = 36 in HRMGPU_F64 (persistent warp-specialised kernel), 3 layers, then a re-sweep."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hrm_b200 import capi  # noqa: E402
from tests import util  # noqa: E402

sc = util.scenes()["3D_superquadrics_cluttered_rabbit"]
rows = sc["arena"] + sc["obstacle"]
semi, R, off, quat = util.body_poses3d(sc["robot"], sc["quats"][:3])
ctx = capi.Context(0)
for n in (10, 23):
    ctx.set_scene3d(1, len(sc["obstacle"]), util.sq17(rows, n), n)
    ctx.set_sweep(sc["lim"], sc["Lx"], sc["Ly"])
    for prec in (capi.F64_STRICT, capi.F64, capi.F32):
        ctx.build_layers(semi, R, off, prec)
        print(n, prec, ctx.dims()["segments"], ctx.single_hit_count())
ctx.set_scene3d(1, len(sc["obstacle"]), util.sq17(rows, 36), 36)
ctx.set_sweep(sc["lim"], 20, 10)
ctx.build_layers(semi, R, off, capi.F64)
print(36, capi.F64, ctx.dims()["segments"], ctx.single_hit_count())
ctx.resweep(sc["lim"], 12, 6)
print("resweep", ctx.dims()["segments"])
ctx.close()
