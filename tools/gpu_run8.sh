mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2m_bench_8gpu.json 2> gpurun_out/r2m_bench_8gpu.err
echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_8gpu.json')); s=d['sharded']; print('ms', s['ms_per_step'], 'floor', s['ingest_floor_ms'], 'x', s['time_over_floor'], s['per_rank_ms_staged_decoded_total'], s['scatter_transport'], 'e2e', d['e2e']['ms_per_step'])"
tail -3 gpurun_out/r2m_bench_8gpu.err
