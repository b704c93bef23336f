timeout 600 python -m pytest tests/test_gpu_fast.py tests/test_gpu_api.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3
for w in cbr320 vbr; do echo "== $w"; timeout 300 python bench.py --no-cpu --no-e2e --workload $w 2>&1 | grep -o 'stage_ms.*'; done
