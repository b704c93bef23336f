# multi-GPU call: tools/gpu_dist.sh <tag> <nproc> [bench]   -> gpurun_out/<tag>_dist_worker.log (+ bench lines)
T=${1:-r2d}; N=${2:-2}; shift; shift
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
ip addr > gpurun_out/${T}_ipaddr.txt 2>&1 || true
NCCL_DEBUG=${NCCL_DEBUG:-WARN} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/dist_worker.py > gpurun_out/${T}_dist_worker.log 2>&1
echo "worker rc=$?"; grep -n "DIST_OK\|Error\|error\|WARN" gpurun_out/${T}_dist_worker.log | head -30
for a in "$@"; do
  if [ "$a" = d2h ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/d2h_floor.py > gpurun_out/${T}_d2h_floor_${N}gpu.json 2>> gpurun_out/${T}_bench_${N}gpu.err
    cat gpurun_out/${T}_d2h_floor_${N}gpu.json
    timeout 300 python tools/d2h_floor.py > gpurun_out/${T}_d2h_floor_1gpu.json 2>> gpurun_out/${T}_bench_${N}gpu.err; cat gpurun_out/${T}_d2h_floor_1gpu.json
  fi
  if [ "$a" = trace ]; then
    P3_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_bench_${N}gpu_trace.json 2> gpurun_out/${T}_trace_${N}gpu.err
    grep -c "pdmp3_read" gpurun_out/${T}_trace_${N}gpu.err; grep "pdmp3_read\|p3 async" gpurun_out/${T}_trace_${N}gpu.err | tail -8
  fi
  if [ "$a" = chunks ]; then
    for ch in ${CHUNKS:-131072 262144 524288}; do
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu --chunk $ch > gpurun_out/${T}_bench_${N}gpu_c$ch.json 2>> gpurun_out/${T}_bench_${N}gpu.err
      python -c "import json,sys; d=json.load(open('gpurun_out/${T}_bench_${N}gpu_c$ch.json')); s=d['sharded']; print('chunk', $ch, 'ms', s['ms_per_step'], 'floor', s['ingest_floor_ms'], 'x', s['time_over_floor'], s['per_rank_ms_staged_decoded_total'])"
    done
  fi
  if [ "$a" = bench ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err
    echo "bench rc=$?"; cut -c1-1500 gpurun_out/${T}_bench_${N}gpu.json; tail -5 gpurun_out/${T}_bench_${N}gpu.err
  fi
done
