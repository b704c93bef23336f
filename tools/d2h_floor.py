"""The host-ingest floor of the e2e path on this box: every rank copies `--mb` MB of device memory into its own pinned host
buffer, all ranks at the same time (python -m torch.distributed.run --nproc-per-node N tools/d2h_floor.py).  Prints one JSON
line: per-rank and aggregate GB/s.  (is the ~87 GB/s aggregate of the 8-GPU e2e run the box's limit?)"""
import argparse, json, os, time
import torch, torch.distributed as dist

ap = argparse.ArgumentParser(); ap.add_argument("--mb", type=int, default=4608); ap.add_argument("--h2d_mb", type=int, default=1045); a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1: dist.init_process_group("gloo")
d = torch.empty(a.mb << 20, dtype=torch.uint8, device="cuda"); h = torch.empty(a.mb << 20, dtype=torch.uint8).pin_memory()
d2 = torch.empty(a.h2d_mb << 20, dtype=torch.uint8, device="cuda"); h2 = torch.empty(a.h2d_mb << 20, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for name, both in (("d2h_only", False), ("d2h_with_h2d", True)):
    ts = []
    for it in range(4):
        if world > 1: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        if both:
            with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    t = torch.tensor([min(ts[1:])], dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[name] = {"seconds_max_over_ranks": float(t.item()), "d2h_GBps_per_rank": (a.mb << 20) / float(t.item()) / 1e9, "d2h_GBps_aggregate": world * (a.mb << 20) / float(t.item()) / 1e9}
if rank == 0: print(json.dumps({"world": world, "mb_per_rank": a.mb, **res}))
if world > 1: dist.destroy_process_group()
