#!/bin/bash
# A/B of the planner-side query kernels on the planner scenes + the parity tests:
#   HRMGPU_NO_BANDS=1         k_transitions3d without band / face boxes and without the triangle-test queue (the plain walk)
#   HRMGPU_NO_FACE_NORMALS=1  k_edges3d without the tabulated face normals
TAG=${1:-r02bands}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_queries.py tests/test_gpu_planner.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
for nb in 0 1; do
  echo "=== HRMGPU_NO_BANDS=$nb HRMGPU_NO_FACE_NORMALS=$nb"
  export HRMGPU_NO_BANDS=$nb HRMGPU_NO_FACE_NORMALS=$nb
  (echo "## cluttered"; HRM_HOST_TIMING=1 python tools/plan_timing.py 3D_superquadrics_cluttered_rabbit 0 10 2>&1 | tail -12
   echo "## maze"; HRM_HOST_TIMING=1 python tools/plan_timing.py 3D_superquadrics_maze_rabbit 0 10 44 20 2>&1 | tail -12) | tee gpurun_out/${TAG}_plan_nb$nb.txt
done
