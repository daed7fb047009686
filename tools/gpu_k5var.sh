#!/bin/bash
# k_transitions3d / k_edges3d device time of one plan() per experiment build (hrm_b200/lib/exp/libhrmgpu_<name>.so, "base" = the
# regular library): the box copy of lib/libhrmgpu.so is swapped, the repository is not touched.
TAG=$1; shift
cp hrm_b200/lib/libhrmgpu.so /tmp/libhrmgpu_base.so
for name in "$@"; do
  if [ "$name" = base ]; then cp /tmp/libhrmgpu_base.so hrm_b200/lib/libhrmgpu.so; else cp hrm_b200/lib/exp/libhrmgpu_$name.so hrm_b200/lib/libhrmgpu.so; fi
  echo "=== $name"
  bash tools/plan_launches.sh ${TAG}_$name 2>&1 | grep -E "launches per|edges3d|transitions3d"
  python tools/plan_timing.py 3D_superquadrics_maze_rabbit 0 10 44 20 2>&1 | grep -E "K5b|plan total" | tail -2
done
cp /tmp/libhrmgpu_base.so hrm_b200/lib/libhrmgpu.so
