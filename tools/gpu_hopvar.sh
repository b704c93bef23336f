mkdir -p gpurun_out
{
for v in default seg8k seg4k; do
  echo "== $v"
  if [ $v = default ]; then L=""; else L="P3_LIB=$PWD/pdmp3_b200/libp3_$v.so"; fi
  env $L python tools/dbg/hop_time.py 4000000 2>&1 | tail -1
  env $L python tools/dbg/hop_time.py 1000000 vbr 2>&1 | tail -1
  env $L python tools/dbg/hop_time.py 125000 2>&1 | tail -1
  env $L python tools/dbg/hop_time.py 32768 2>&1 | tail -1
  env $L timeout 600 python -m pytest tests/test_gpu_hop.py -x -q -m gpu 2>&1 | tail -1
done
} > gpurun_out/r2q_hop_segments.log 2>&1
cat gpurun_out/r2q_hop_segments.log
