#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python tools/stencil_ab.py both gpurun_out/stencil_ab4.txt > gpurun_out/stencil_ab4.log 2>&1; echo "ab rc=$?"
grep -v "^    at" gpurun_out/stencil_ab4.txt | cut -c1-250
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -25 > gpurun_out/r2b_pytest4.txt; tail -3 gpurun_out/r2b_pytest4.txt
