#!/bin/bash
# Regenerates the measurements behind profiles/ in one gpurun call (outputs under gpurun_out/; copy/summarise afterwards with
# tools/summarize_profile.py and tools/source_hotspots.py).  Never run two of these at once; numbers printed under ncu are not bench values.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
python bench.py --precision f32 --no-cpu-baseline > gpurun_out/b_f32.json 2>/dev/null; cut -c1-200 gpurun_out/b_f32.json
python bench.py --precision f64_strict --no-cpu-baseline > gpurun_out/b_f64_strict.json 2>/dev/null; cut -c1-200 gpurun_out/b_f64_strict.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference.json
(echo "## cluttered/rabbit, 11x5 lines"; HRM_HOST_TIMING=1 python tools/plan_timing.py 3D_superquadrics_cluttered_rabbit 0 10 2>&1; echo "## maze/rabbit, 44x20 lines"; HRM_HOST_TIMING=1 python tools/plan_timing.py 3D_superquadrics_maze_rabbit 0 10 44 20 2>&1) > gpurun_out/plan_stages.txt; tail -3 gpurun_out/plan_stages.txt
python tools/k1_only.py f64 128 > gpurun_out/k1_only.txt 2>&1; python tools/k1_only.py f32 128 >> gpurun_out/k1_only.txt 2>&1; cat gpurun_out/k1_only.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 3 --layers 16 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_minksweep3d -s 2 -c 1 -o gpurun_out/r01_minksweep3d_f64 -f python tools/prof_layers.py f64 16 3 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
python tools/phase_times.py f64 16 > gpurun_out/phase_times_f64.txt 2>&1
python tools/phase_times.py f32 16 > gpurun_out/phase_times_f32.txt 2>&1
ls -la gpurun_out | tail -8
