#!/bin/bash
# Phase clocks of the fused kernel in its variants (profiling aid).  Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-prof}
for v in "" "HRMGPU_WS_REBAL=1" "HRMGPU_CLASSIC=1"; do
  echo "=== variant: ${v:-ws 80/80}" >> gpurun_out/${TAG}_phases.txt
  env $v timeout 600 python tools/phase_times.py f64 32 >> gpurun_out/${TAG}_phases.txt 2>&1
done
cat gpurun_out/${TAG}_phases.txt
