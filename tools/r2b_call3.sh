#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/stencil_ab.py both gpurun_out/stencil_ab3.txt > gpurun_out/stencil_ab3.log 2>&1; echo "ab rc=$?"
grep -v "^DIFF\|^    at" gpurun_out/stencil_ab3.txt | cut -c1-330
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -25 > gpurun_out/r2b_pytest3.txt; tail -3 gpurun_out/r2b_pytest3.txt
