# A/B of kernel variants on one box: tools/gpu_exp.sh  (variant libraries built by tools/build_variant.sh, selected through P3_LIB)
mkdir -p gpurun_out
( P3_LIB=$PWD/pdmp3_b200/libp3_k1sa.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3 ) > gpurun_out/exp_tests.log 2>&1
b() { echo "== $*"; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --workload ${WL:-cbr320} 2>&1 | grep -o '"ms_per_step[^,]*\|stage_ms.*' | tr '\n' ' '; echo; }
{
b P3_X=1
b P3_LIB=$PWD/pdmp3_b200/libp3_k1lean.so
b P3_LIB=$PWD/pdmp3_b200/libp3_k1sa.so
WL=vbr b P3_X=1
WL=vbr b P3_LIB=$PWD/pdmp3_b200/libp3_k1sa.so
} > gpurun_out/exp_bench.log 2>&1
cat gpurun_out/exp_tests.log gpurun_out/exp_bench.log
