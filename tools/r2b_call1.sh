#!/bin/bash
# round 2, second session, GPU call 1: the two stencil cuts side by side, the GPU tests, the default bench
set -x
mkdir -p gpurun_out
timeout 300 python tools/stencil_ab.py both gpurun_out/stencil_ab.txt > gpurun_out/stencil_ab.log 2>&1; echo "ab rc=$?"
tail -5 gpurun_out/stencil_ab.txt
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -25 > gpurun_out/r2b_pytest.txt; tail -3 gpurun_out/r2b_pytest.txt
timeout 420 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err
