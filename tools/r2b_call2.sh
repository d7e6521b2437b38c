#!/bin/bash
# ncu --set full with source correlation of both stencil cuts on a 64-map batch
set -x
mkdir -p gpurun_out
timeout 240 ncu --set full --import-source on --clock-control none -k regex:lsdb_stencil -c 2 -f -o gpurun_out/prof_st2 python tools/run_batch.py 1000 64 > gpurun_out/prof_st2.log 2>&1; echo rc=$?
LSDB_STENCIL=1 timeout 240 ncu --set full --import-source on --clock-control none -k regex:lsdb_stencil -c 1 -f -o gpurun_out/prof_st1 python tools/run_batch.py 1000 64 > gpurun_out/prof_st1.log 2>&1; echo rc=$?
ls -la gpurun_out/prof_st*.ncu-rep; tail -2 gpurun_out/prof_st2.log
