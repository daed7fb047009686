#!/bin/bash
# Experiment builds side by side: plain kernel timing + phase clocks of every hrm_b200/lib/exp/libhrmgpu_<name>.so given.
#   tools/gpu_exp.sh <tag> <name> [<name> ...]      ("base" = the regular library)
mkdir -p gpurun_out
TAG=$1; shift
OUT=gpurun_out/${TAG}_exp.txt
rm -f $OUT
for name in "$@"; do
  if [ "$name" = base ]; then unset HRMGPU_LIB; else export HRMGPU_LIB=$PWD/hrm_b200/lib/exp/libhrmgpu_$name.so; fi
  echo "=== $name: plain timing" >> $OUT
  HRM_CASE_REPS=${HRM_CASE_REPS:-8} timeout 300 python tools/ncu_case.py f64 125 2>&1 | tail -1 >> $OUT
  if [ -z "$NO_PHASES" ]; then
    echo "=== $name: phase clocks" >> $OUT
    timeout 300 python tools/phase_times.py f64 64 2>&1 | tail -9 >> $OUT
  fi
done
cat $OUT
