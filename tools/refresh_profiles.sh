#!/bin/bash
# Regenerates the measurements behind profiles/ in one gpurun call (outputs under gpurun_out/; copy/summarise afterwards with
# tools/ncu_brief.py and tools/source_hotspots.py).  Never run two of these at once; numbers printed under ncu are not bench values.
TAG=${1:-r02}
mkdir -p gpurun_out
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${TAG}_bench_default_f64.json 2> gpurun_out/${TAG}_bench_default.err; cut -c1-300 gpurun_out/${TAG}_bench_default_f64.json
timeout 600 python bench.py --precision f32 --no-cpu-baseline --no-freespace > gpurun_out/${TAG}_bench_f32.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_f32.json
timeout 600 python bench.py --precision f64_strict --no-cpu-baseline --no-freespace --no-parity > gpurun_out/${TAG}_bench_f64_strict.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_f64_strict.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_bench_reference.json
(echo "## cluttered/rabbit, 11x5 lines"; HRM_HOST_TIMING=1 python tools/plan_timing.py 3D_superquadrics_cluttered_rabbit 0 10 2>&1; echo "## maze/rabbit, 44x20 lines"; HRM_HOST_TIMING=1 python tools/plan_timing.py 3D_superquadrics_maze_rabbit 0 10 44 20 2>&1) > gpurun_out/${TAG}_plan_stages.txt; tail -3 gpurun_out/${TAG}_plan_stages.txt
bash tools/plan_launches.sh ${TAG} > gpurun_out/${TAG}_plan_kernels.txt 2>&1; cat gpurun_out/${TAG}_plan_kernels.txt
HRM_HOST_TIMING=1 python tools/prob_timing.py 3 > gpurun_out/${TAG}_prob_stages.txt 2>&1; tail -2 gpurun_out/${TAG}_prob_stages.txt
python tools/k1_only.py f64 125 > gpurun_out/${TAG}_k1_only.txt 2>&1; cat gpurun_out/${TAG}_k1_only.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --layers 16 --no-cpu-baseline --no-parity --no-freespace > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_minksweep3d -s 2 -c 1 -o gpurun_out/${TAG}_minksweep3d_ws_f64 -f python tools/prof_layers.py f64 16 3 > gpurun_out/${TAG}_ncu2.log 2>&1; tail -2 gpurun_out/${TAG}_ncu2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_line_segments -s 2 -c 1 -o gpurun_out/${TAG}_line_segments -f python tools/prof_layers.py f64 16 3 > gpurun_out/${TAG}_ncu3.log 2>&1; tail -2 gpurun_out/${TAG}_ncu3.log
timeout 300 python tools/phase_times.py f64 64 > gpurun_out/${TAG}_phase_times_f64.txt 2>&1
HRMGPU_CLASSIC=1 timeout 300 python tools/phase_times.py f64 16 > gpurun_out/${TAG}_phase_times_f64_classic.txt 2>&1
ls -la gpurun_out | tail -20
