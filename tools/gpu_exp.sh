# A/B of kernel variants on one box: tools/gpu_exp.sh  (variant libraries built by tools/build_variant.sh, selected through P3_LIB)
mkdir -p gpurun_out
b() { echo "== $*"; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --workload ${WL:-cbr320} 2>&1 | grep -o '"ms_per_step[^,]*\|stage_ms.*' | tr '\n' ' '; echo; }
{
b P3_X=1
b P3_LIB=$PWD/pdmp3_b200/libp3_w16.so
b P3_LIB=$PWD/pdmp3_b200/libp3_w16b.so
b P3_LIB=$PWD/pdmp3_b200/libp3_nb1.so
WL=vbr b P3_X=1
WL=vbr b P3_LIB=$PWD/pdmp3_b200/libp3_w16b.so
} > gpurun_out/exp_bench.log 2>&1
cat gpurun_out/exp_bench.log
