set -x
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/s6_tests.log 2>&1; tail -5 gpurun_out/s6_tests.log
timeout 600 python bench.py --no-cpu > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err; cat gpurun_out/s6_bench.json; tail -5 gpurun_out/s6_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_huffman -s 2 -c 1 -f -o gpurun_out/s6_huff python bench.py --frames 125000 --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
