( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s10_tests.log 2>&1; tail -4 gpurun_out/s10_tests.log
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err; cat gpurun_out/s10_bench.json | grep -o '"ms_per_step": [0-9.]*\|stage_ms.*'; tail -3 gpurun_out/s10_bench.err
