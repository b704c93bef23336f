# round-2 (third session) validation call: full GPU suite, smoke, bench lines, reference arm with the CPU-baseline variants,
# A/B of the K1 pair2 variant (quick timing + its parity subset), memcheck of the overlapped pipeline
T=${1:-r3b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > gpurun_out/${T}_smoke.log 2>&1; cat gpurun_out/${T}_smoke.log
timeout 400 python bench.py > gpurun_out/${T}_bench_fast.json 2> gpurun_out/${T}_bench.err
timeout 200 python bench.py --no-cpu --workload vbr > gpurun_out/${T}_bench_vbr.json 2>> gpurun_out/${T}_bench.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
for f in fast vbr reference; do cut -c1-300 gpurun_out/${T}_bench_$f.json; echo; done; tail -3 gpurun_out/${T}_bench.err
for v in $VARIANTS; do
  echo "== $v" >> gpurun_out/${T}_variants.log
  P3_LIB=$PWD/pdmp3_b200/libp3_$v.so timeout 200 python tools/dbg/overlap_sweep.py 1000000 quick >> gpurun_out/${T}_variants.log 2>> gpurun_out/${T}_variants.err
  P3_LIB=$PWD/pdmp3_b200/libp3_$v.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast.py tests/test_gpu_iso.py tests/test_gpu_fuzz.py -x -q -m gpu -k "not one_million and not bench_block" 2>&1 | tail -2 >> gpurun_out/${T}_variants.log
done
cat gpurun_out/${T}_variants.log
( P3_OVERLAP=128,1 P3_SAN_FRAMES=300 timeout 150 compute-sanitizer --tool memcheck --print-limit 20 python tools/dbg/sanitize.py cfg4 cfg3 2>&1 | tail -8 ) > gpurun_out/${T}_sanitizer_overlap_memcheck.log 2>&1; tail -3 gpurun_out/${T}_sanitizer_overlap_memcheck.log
ls gpurun_out | grep ${T}_ | tr '\n' ' '
