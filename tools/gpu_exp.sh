# A/B of kernel variants on one box (one gpurun call): build the variants first with tools/build_variant.sh NAME FLAGS..., e.g.
#   P3_SRC=/tmp/head bash tools/build_variant.sh base          (an export of an older commit: git archive HEAD pdmp3_b200/csrc include | tar -x -C /tmp/head)
#   bash tools/build_variant.sh w16 -DSW_WPB=8 -DSW_MINB=2 -DSW_NBUF=1
# then: gpurun -- 'VARIANTS="base w16" bash tools/gpu_exp.sh'   -> gpurun_out/exp_bench.log (stage times of the default library and of every variant)
mkdir -p gpurun_out
b() { echo "== $*"; env "$@" timeout 300 python bench.py --no-cpu --no-e2e --workload ${WL:-cbr320} 2>&1 | grep -o '"ms_per_step[^,]*\|stage_ms.*' | tr '\n' ' '; echo; }
{
for wl in cbr320 vbr; do
  WL=$wl b P3_X=default
  for v in $VARIANTS; do WL=$wl b P3_LIB=$PWD/pdmp3_b200/libp3_$v.so; done
done
} > gpurun_out/exp_bench.log 2>&1
cat gpurun_out/exp_bench.log
