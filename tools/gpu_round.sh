set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r1b_tests.log 2>&1; tail -4 gpurun_out/r1b_tests.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1b_bench_reference.json 2> gpurun_out/r1b_ref.err; cat gpurun_out/r1b_bench_reference.json
timeout 600 python bench.py > gpurun_out/r1b_bench_fast.json 2> gpurun_out/r1b_bench.err; cat gpurun_out/r1b_bench_fast.json; tail -3 gpurun_out/r1b_bench.err
timeout 600 python bench.py --workload vbr --no-cpu > gpurun_out/r1b_bench_vbr.json 2>/dev/null; cat gpurun_out/r1b_bench_vbr.json
timeout 600 python bench.py --mode exact --no-cpu --no-e2e > gpurun_out/r1b_bench_exact.json 2>/dev/null; cat gpurun_out/r1b_bench_exact.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
for k in k_synth_warp k_huffman k_compact; do timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r1b_$k python bench.py --frames 125000 --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1; done
ls -la gpurun_out | tail -12
