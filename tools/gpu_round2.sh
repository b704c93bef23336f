# Round-2 GPU box call: tools/gpu_round2.sh <tag> [steps...]   steps: tests smoke bench san ncu   (default: all)
T=${1:-r2x}; shift
STEPS=${*:-tests smoke bench san ncu}
mkdir -p gpurun_out
has() { case " $STEPS " in *" $1 "*) return 0;; esac; return 1; }
if has tests; then timeout 1500 python -m pytest ${PYTEST_ARGS:-tests} -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -15 gpurun_out/${T}_tests.log; fi
if has smoke; then ( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > gpurun_out/${T}_smoke.log 2>&1; cat gpurun_out/${T}_smoke.log; fi
if has bench; then
  timeout 600 python bench.py > gpurun_out/${T}_bench_fast.json 2> gpurun_out/${T}_bench.err
  timeout 300 python bench.py --no-cpu --workload vbr > gpurun_out/${T}_bench_vbr.json 2>> gpurun_out/${T}_bench.err
  timeout 300 python bench.py --no-cpu --workload xr > gpurun_out/${T}_bench_xr.json 2>> gpurun_out/${T}_bench.err
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err
  for f in fast vbr xr; do cut -c1-700 gpurun_out/${T}_bench_$f.json; echo; done; tail -3 gpurun_out/${T}_bench.err
fi
if has san; then
  for tool in memcheck racecheck; do
    ( P3_SAN_FRAMES=300 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/dbg/sanitize.py cfg4 2>&1 | tail -25 ) > gpurun_out/${T}_sanitizer_$tool.log 2>&1
    tail -4 gpurun_out/${T}_sanitizer_$tool.log
  done
fi
if has ncu; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches_fast.csv python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 > /dev/null 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${T}_launches_vbr.csv python bench.py --no-cpu --no-e2e --workload vbr --steps 2 --warmup 1 > /dev/null 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_synth_warp_same\$ -s 1 -c 1 -o gpurun_out/${T}_vbr_k_synth_warp_same -f python bench.py --no-cpu --no-e2e --workload vbr --frames 125000 --steps 1 --warmup 1 > /dev/null 2>&1
  for k in ${NCU_KERNELS:-k_synth_warp_lean k_huffman}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 1 -c 1 -o gpurun_out/${T}_$k -f python bench.py --no-cpu --no-e2e --frames 125000 --steps 1 --warmup 1 > /dev/null 2>&1
  done
  ls -la gpurun_out | grep ${T}
fi
