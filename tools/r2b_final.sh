#!/bin/bash
# round 2, second session, final GPU call: sanitizer on the new stencil kernels, ncu --set full of one stencil launch (traffic),
# the default bench, the launch list of the bench command
set -x
mkdir -p gpurun_out
{
  echo "# compute-sanitizer (CUDA 12.9) on a B200, kernels filtered to lsdb_stencil*, __graft_entry__.smoke()"
  echo "## --tool memcheck"
  timeout 120 compute-sanitizer --tool memcheck --kernel-name kns=lsdb_stencil python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^Score" | tail -6
  echo "## --tool racecheck"
  timeout 120 compute-sanitizer --tool racecheck --kernel-name kns=lsdb_stencil python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^Score" | tail -6
} > gpurun_out/r2b_sanitizer.txt 2>&1
tail -4 gpurun_out/r2b_sanitizer.txt
timeout 200 ncu --set full --import-source on --clock-control none -k regex:lsdb_stencil -c 2 -f -o gpurun_out/prof_st2_final python tools/run_batch.py 1000 64 > gpurun_out/prof_st2_final.log 2>&1; echo rc=$?
python tools/stencil_traffic.py gpurun_out/prof_st2_final.ncu-rep $((64*4096*4096)) profiles/r2_stencil_traffic.json && cp profiles/r2_stencil_traffic.json gpurun_out/r2b_stencil_traffic.json
timeout 420 python bench.py > gpurun_out/r2b_final_bench.json 2> gpurun_out/r2b_final_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2b_final_bench.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --parity-maps 0 > gpurun_out/r2b_launches_bench.log 2>&1; echo "launch list rc=$?"
