timeout 300 python -m pytest tests/test_gpu_rdp.py tests/test_gpu_dropin.py -x -q 2>&1 | tail -4
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_rdp.py -x -q -k "ragged or non_default or golden" 2>&1 | grep -v "^Score" | tail -4
timeout 200 python tools/fscan_probe.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:lsdb_fscan -c 4 -o gpurun_out/prof_fscan_r1z -f python tools/fscan_probe.py > gpurun_out/ncu_fscan.log 2>&1; tail -2 gpurun_out/ncu_fscan.log
