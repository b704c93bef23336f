timeout 600 python -m pytest tests/test_gpu_fast.py -x -q -k "vbr_mixed" 2>&1 | tail -3
timeout 600 python bench.py --no-cpu --workload vbr > gpurun_out/s12_bench_vbr.json 2> gpurun_out/s12_bench_vbr.err; cat gpurun_out/s12_bench_vbr.json | grep -o '"ms_per_step": [0-9.]*\|stage_ms.*\|"value": [0-9.]*'; tail -3 gpurun_out/s12_bench_vbr.err
