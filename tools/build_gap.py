"""Back-to-back builds: time per build against the fused kernel's own time (where does a step spend what is not the kernel?).
    python tools/build_gap.py [layers]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hrm_b200 import capi, workload as wl  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 125
dev = torch.device("cuda", 0)
ctx = capi.Context(0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
rows = wl.scene_rows12()
ctx.set_scene3d(1, wl.N_OBS, wl.scene_rows17(rows), wl.N_GRID)
ctx.set_sweep(wl.LIM, wl.LX, wl.LY)
nb = 24
R = torch.tensor(wl.body_rotations(wl.orientations(nb * layers)), dtype=torch.float64, device=dev)
off = torch.zeros((nb * layers, 1, 3), dtype=torch.float64, device=dev)
semi = torch.tensor(np.array(wl.BODY_SEMI), dtype=torch.float64, device=dev)


def build(i):
    a = i * layers
    ctx.build_layers_dev(layers, 1, semi.data_ptr(), R[a:a + layers].data_ptr(), off[a:a + layers].data_ptr(), capi.F64)


for i in range(4):
    build(i)
ctx.synchronize()
for B in (1, 2, 4, 8, 16):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(B):
        build(4 + i)
    ctx.join()
    e1.record(stream)
    torch.cuda.synchronize()
    print("%2d builds back to back: %.3f ms per build (fused kernel of the last one %.3f ms, its free-segment passes %.3f ms)"
          % (B, e0.elapsed_time(e1) / B, ctx.last_minksweep_ms(), ctx.last_segments_ms()))
ctx.close()
