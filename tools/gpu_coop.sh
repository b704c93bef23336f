# warp-cooperative Huffman trial (tools/build_variant.sh coop -DK1_COOP_TRIAL): parity tests on the variant, then its K1 time next to the product's
mkdir -p gpurun_out
P3_LIB=$PWD/pdmp3_b200/libp3_coop.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stage_taps or empty_parts" > gpurun_out/r2p_coop_tests.log 2>&1; tail -4 gpurun_out/r2p_coop_tests.log
{
for wl in cbr320 vbr; do
  for lib in default coop sortp23; do
    echo "== $wl $lib"
    if [ $lib = default ]; then timeout 300 python bench.py --no-cpu --no-e2e --workload $wl 2>&1 | grep -o '"ms_per_step[^,]*\|stage_ms[^}]*}'
    else P3_LIB=$PWD/pdmp3_b200/libp3_$lib.so timeout 300 python bench.py --no-cpu --no-e2e --workload $wl 2>&1 | grep -o '"ms_per_step[^,]*\|stage_ms[^}]*}'; fi
  done
done
} > gpurun_out/r2p_coop_bench.log 2>&1
cat gpurun_out/r2p_coop_bench.log
