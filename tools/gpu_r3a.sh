# round-2 (third session) experiment call: K1 / synthesis overlap sweep, frames-per-warp sweep, K1 LUT-in-global variants, the new overlap test
T=${1:-r3a}
mkdir -p gpurun_out
timeout 400 python tools/dbg/overlap_sweep.py > gpurun_out/${T}_overlap_sweep.log 2> gpurun_out/${T}_overlap_sweep.err; tail -3 gpurun_out/${T}_overlap_sweep.err
grep -c . gpurun_out/${T}_overlap_sweep.log; grep "sequential\"\|best_overlap" gpurun_out/${T}_overlap_sweep.log
for v in lutg lutg64; do
  echo "== $v" >> gpurun_out/${T}_variants.log
  P3_LIB=$PWD/pdmp3_b200/libp3_$v.so timeout 200 python tools/dbg/overlap_sweep.py 1000000 quick >> gpurun_out/${T}_variants.log 2>> gpurun_out/${T}_variants.err
  P3_LIB=$PWD/pdmp3_b200/libp3_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast.py -x -q -m gpu -k "stage_taps or warp_kernel or overlapped" 2>&1 | tail -2 >> gpurun_out/${T}_variants.log
done
cat gpurun_out/${T}_variants.log
timeout 300 python -m pytest tests/test_gpu_fast.py -x -q -m gpu -k "overlapped or partition or content_classes" > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
