# ncu launch list of bench.py on the final tree (one GPU)
T=${1:-r3d}
mkdir -p gpurun_out
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches_fast.csv python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 > /dev/null 2>&1
grep -c . gpurun_out/${T}_launches_fast.csv; tail -8 gpurun_out/${T}_launches_fast.csv | cut -c1-200
