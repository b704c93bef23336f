( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s8_tests.log 2>&1; tail -6 gpurun_out/s8_tests.log
timeout 600 python bench.py --no-cpu > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err; cat gpurun_out/s8_bench.json | grep -o '"ms_per_step": [0-9.]*\|stage_ms.*\|"e2e": {"value": [0-9.]*\|"ms_per_step": [0-9.]*, "api'; tail -3 gpurun_out/s8_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_huffman -s 2 -c 1 -f -o gpurun_out/s8_huff python bench.py --frames 125000 --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()"
