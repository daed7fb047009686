#!/bin/bash
# One GPU-box pass: GPU parity tests, then the default bench line.  Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-run}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
