for ch in 16384 32768 65536 131072 262144; do echo "== two streams, chunk $ch"; P3_PINGPONG=1 P3_CHUNK=$ch timeout 300 python bench.py --no-cpu --no-e2e 2>&1 | grep -o '"ms_per_step": [0-9.]*'; done
echo "== default"; timeout 300 python bench.py --no-cpu --no-e2e 2>&1 | grep -o '"ms_per_step": [0-9.]*'
P3_PINGPONG=1 P3_CHUNK=65536 timeout 600 python -m pytest tests/test_gpu_fast.py -x -q -k "million or partition" 2>&1 | tail -2
