# A/B of synthesis-occupancy variants (register cap instead of a CTA count): quick kernel timings only
T=${1:-r3e}
mkdir -p gpurun_out
for v in $VARIANTS; do
  echo "== $v" >> gpurun_out/${T}_variants.log
  P3_LIB=$PWD/pdmp3_b200/libp3_$v.so timeout 60 python tools/dbg/overlap_sweep.py 1000000 quick >> gpurun_out/${T}_variants.log 2>> gpurun_out/${T}_variants.err
done
cat gpurun_out/${T}_variants.log; tail -2 gpurun_out/${T}_variants.err
