# final round-2 validation on one B200: tests, smoke, bench lines, variant A/B, sanitizers on the new kernels, ncu launch lists + captures
T=${1:-r2z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/${T}_smoke.log 2>&1; cat gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_fast.json 2> gpurun_out/${T}_bench.err
timeout 300 python bench.py --no-cpu --workload vbr > gpurun_out/${T}_bench_vbr.json 2>> gpurun_out/${T}_bench.err
timeout 300 python bench.py --no-cpu --workload xr > gpurun_out/${T}_bench_xr.json 2>> gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
for f in fast vbr xr reference; do cut -c1-260 gpurun_out/${T}_bench_$f.json; echo; done; tail -3 gpurun_out/${T}_bench.err
{ for wl in vbr cbr320; do echo "== $wl ctasync"; P3_LIB=$PWD/pdmp3_b200/libp3_ctasync.so timeout 300 python bench.py --no-cpu --no-e2e --workload $wl 2>&1 | grep -o '"ms_per_step[^,]*\|stage_ms[^}]*}'; done; } > gpurun_out/${T}_ctasync_bench.log 2>&1; cat gpurun_out/${T}_ctasync_bench.log
P3_LIB=$PWD/pdmp3_b200/libp3_ctasync.so timeout 300 python -m pytest tests/test_gpu_fast.py -x -q -m gpu -k "content_classes or warp_kernel" > gpurun_out/${T}_ctasync_tests.log 2>&1; tail -2 gpurun_out/${T}_ctasync_tests.log
for tool in memcheck racecheck; do
  ( timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/dbg/sanitize_hop.py 2>&1 | tail -12 ) > gpurun_out/${T}_sanitizer_hop_$tool.log 2>&1; tail -3 gpurun_out/${T}_sanitizer_hop_$tool.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches_fast.csv python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches_vbr.csv python bench.py --no-cpu --no-e2e --workload vbr --steps 2 --warmup 1 > /dev/null 2>&1
for k in k_synth_warp_lean k_huffman k_compact k_hop_spec; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 1 -c 1 -o gpurun_out/${T}_$k -f python bench.py --no-cpu --no-e2e --frames 125000 --steps 1 --warmup 1 > /dev/null 2>&1
done
ls gpurun_out | grep ${T}_ | tr '\n' ' '
