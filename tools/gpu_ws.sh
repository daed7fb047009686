#!/bin/bash
# WS kernel iteration: its tests, plain kernel timing of the variants (no instrumentation), then phase clocks.
mkdir -p gpurun_out
TAG=${1:-ws}
timeout 900 python -m pytest tests/test_gpu_ws.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
rm -f gpurun_out/${TAG}_phases.txt
for v in ${TIME_VARIANTS:-"HRMGPU_WS_VARIANT=1" "HRMGPU_WS_VARIANT=0" "HRMGPU_WS_GENERIC=1" "HRMGPU_CLASSIC=1"}; do
  echo "=== plain timing, variant: ${v}" >> gpurun_out/${TAG}_phases.txt
  env $v timeout 600 python tools/ncu_case.py f64 125 >> gpurun_out/${TAG}_phases.txt 2>&1
done
for v in ${PHASE_VARIANTS:-1}; do
  echo "=== phase clocks, variant: ${v}" >> gpurun_out/${TAG}_phases.txt
  env HRMGPU_WS_VARIANT=$v timeout 600 python tools/phase_times.py f64 64 >> gpurun_out/${TAG}_phases.txt 2>&1
done
cat gpurun_out/${TAG}_phases.txt
