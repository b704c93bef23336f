mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_iso.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2t_tests.log 2>&1; tail -3 gpurun_out/r2t_tests.log
{ for wl in vbr cbr320; do echo "== $wl outlined"; timeout 300 python bench.py --no-cpu --no-e2e --workload $wl 2>&1 | grep -o '"ms_per_step[^,]*\|stage_ms[^}]*}'; done; } > gpurun_out/r2t_outline_bench.log 2>&1; cat gpurun_out/r2t_outline_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_synth_warp_same\$ -s 1 -c 1 -o gpurun_out/r2t_vbr_k_synth_warp_same -f python bench.py --no-cpu --no-e2e --workload vbr --frames 125000 --steps 1 --warmup 1 > /dev/null 2>&1
