( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s11_tests.log 2>&1; tail -4 gpurun_out/s11_tests.log
timeout 600 python bench.py --no-cpu --no-e2e > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err; cat gpurun_out/s11_bench.json | grep -o '"ms_per_step": [0-9.]*\|stage_ms.*'; tail -3 gpurun_out/s11_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_synth_warp -s 2 -c 1 -f -o gpurun_out/s11_synthw python bench.py --frames 125000 --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
