# One GPU-box round: tools/gpu_round.sh [tag]  -> gpurun_out/<tag>_*  (tests, bench lines, ncu launch list, ncu --set full of the kernels)
T=${1:-rX}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) > gpurun_out/${T}_tests.log 2>&1
timeout 600 python bench.py > gpurun_out/${T}_bench_fast.json 2> gpurun_out/${T}_bench_fast.err
timeout 300 python bench.py --no-cpu --workload vbr > gpurun_out/${T}_bench_vbr.json 2>> gpurun_out/${T}_bench_fast.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench_fast.err
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_launches_fast.csv python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 > /dev/null 2>&1
for k in k_synth_warp_lean k_huffman; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 1 -c 1 -o gpurun_out/${T}_$k -f python bench.py --no-cpu --no-e2e --frames 125000 --steps 1 --warmup 1 > /dev/null 2>&1
done
fi
cat gpurun_out/${T}_tests.log; cut -c1-400 gpurun_out/${T}_bench_fast.json; tail -3 gpurun_out/${T}_bench_fast.err
