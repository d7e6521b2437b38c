timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^Score" | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "^Score" | tail -3
timeout 900 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 2>&1 | tail -1 | cut -c1-700
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_final_check.json 2>gpurun_out/bench_final_check.err; tail -c 1500 gpurun_out/bench_final_check.json; tail -3 gpurun_out/bench_final_check.err
