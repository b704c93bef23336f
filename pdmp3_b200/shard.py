"""Frame sharding of one MP3 stream over the GPUs of a box.

Frames are independent units up to a fixed halo, so the path shards with NO data-path collective:
each rank decodes a contiguous run of frames plus a short warm-up in front of it (the frame before
the run primes the IMDCT overlap and the 15-slot polyphase history; frames before that only supply
bit-reservoir bytes) and drops the warm-up PCM.  The only communication is what BASELINE.json's
north_star names: scatter of the compressed bytes from rank 0 and gather of the PCM to rank 0
(torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import numpy as np


def plan_shards(frames, world):
    """frames: structured array of p3_frame records of the WHOLE stream (parser output).
    Returns one dict per rank: first/last frame (PCM owned), warm-up frame count, byte range to ship."""
    n = len(frames)
    plans = []
    for r in range(world):
        a, b = n * r // world, n * (r + 1) // world
        w = a
        if a > 0 and b > a:
            w = a - 1                                        # the frame that must decode correctly
            need = int(frames["main_pos"][w]) - int(frames["main_begin"][w])
            while w > 0 and int(frames["main_pos"][w]) > need:
                w -= 1                                       # earlier frames: reservoir bytes only
        lo = 0 if w == 0 else int(frames["main_off"][w - 1]) + int(frames["main_size"][w - 1])
        hi = lo if b <= a else int(frames["main_off"][b - 1]) + int(frames["main_size"][b - 1])
        plans.append(dict(rank=r, first=a, last=b, warmup=a - w, byte_lo=lo, byte_hi=hi))
    return plans


def decode_sharded(stream, parse, decode, rank, world, device="cpu", group=None):
    """Rank 0 holds `stream` (np.uint8) and parses it; every rank receives its byte range, decodes it
    with `decode(bytes_np, warmup) -> int16 array [frames,1152,nch]`, and rank 0 gets the PCM of the
    whole stream back (None elsewhere).  `parse(stream) -> frames structured array`."""
    import torch, torch.distributed as dist
    if world == 1:
        return decode(stream, 0)
    meta = [None]
    if rank == 0:
        plans = plan_shards(parse(stream), world)
        meta = [plans]
    dist.broadcast_object_list(meta, src=0, group=group)
    plans = meta[0]
    me = plans[rank]
    # ---- scatter of compressed bytes (point to point: the ranges overlap by the warm-up) ----
    if rank == 0:
        full = torch.from_numpy(np.ascontiguousarray(stream)).to(device)
        reqs = [dist.isend(full[p["byte_lo"]:p["byte_hi"]].contiguous(), dst=p["rank"], group=group) for p in plans[1:]]
        mine = full[me["byte_lo"]:me["byte_hi"]]
        for q in reqs: q.wait()
    else:
        mine = torch.empty(me["byte_hi"] - me["byte_lo"], dtype=torch.uint8, device=device)
        dist.recv(mine, src=0, group=group)
    pcm = decode(mine.cpu().numpy(), me["warmup"])
    assert pcm.shape[0] == me["last"] - me["first"], (pcm.shape, me)
    # ---- gather of PCM to rank 0 ----
    t = torch.from_numpy(np.ascontiguousarray(pcm)).to(device)
    if rank == 0:
        parts = [t] + [torch.empty((p["last"] - p["first"],) + tuple(t.shape[1:]), dtype=t.dtype, device=device) for p in plans[1:]]
        reqs = [dist.irecv(parts[p["rank"]], src=p["rank"], group=group) for p in plans[1:]]
        for q in reqs: q.wait()
        return torch.cat(parts).cpu().numpy()
    dist.send(t, dst=0, group=group)
    return None
