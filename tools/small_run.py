"""Tiny deterministic run of every precision mode (for compute-sanitizer): bundled cluttered map, n=10, 3 layers."""
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
ctx.close()
