mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/r1c_smoke.log 2>&1; tail -2 gpurun_out/r1c_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k k_synth_warp -s 1 -c 1 -o gpurun_out/r1c_vbr_k_synth_warp -f python bench.py --no-cpu --no-e2e --workload vbr --frames 125000 --steps 1 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
