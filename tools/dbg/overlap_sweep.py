#!/usr/bin/env python3
"""K1 / synthesis overlap sweep (p3_ctx_set_overlap) and frames-per-warp sweep on one GPU, both bench workloads:
   python tools/dbg/overlap_sweep.py [frames] > gpurun_out/<tag>_overlap_sweep.log
Every configuration is timed with p3_batch_time (CUDA events on the context stream, 5 steps after 2 warm-up runs) and its PCM is
compared, on the device, with the sequential path's PCM of the same batch."""
import os, sys, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch, p3synth, pdmp3_b200

NF = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
QUICK = len(sys.argv) > 2 and sys.argv[2] == "quick"       # only the sequential path (A/B of a variant library through P3_LIB)
BLOCK = 15625
CFGS = {"cbr320": dict(bitrate_index=14, mode=1, mode_ext=2, blocks=0),
        "vbr": dict(bitrate_index=0, mode=1, mode_ext=-1, blocks=1, overrun_pm=30)}
W2, W3 = 148 * 2 * 4 * 32, 148 * 3 * 4 * 32          # frames in one wave of the synthesis kernel at 2 / 3 CTAs per SM


def pcm_dev(ctx, n_frames):
    ptr = pdmp3_b200.lib().p3_batch_pcm_device(ctx.h, None)
    class W: pass
    w = W(); w.__cuda_array_interface__ = {"shape": (n_frames * 1152,), "typestr": "<i4", "data": (ptr, False), "version": 2}
    return torch.as_tensor(w, device="cuda")


def measure(ctx, stream, n_iter=5):
    info = ctx.upload_raw(stream, lookahead=0); ctx.sync()
    ctx.run(); ctx.run(); ctx.sync()
    ms, st = ctx.time(n_iter)
    return info, ms, st


for wl, cfg in CFGS.items():
    blk, _ = p3synth.synth(min(BLOCK, NF), seed=1, **cfg)
    stream = np.tile(blk, (NF + BLOCK - 1) // BLOCK) if NF > BLOCK else blk
    ctx = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST)
    info, ms0, st0 = measure(ctx, stream)
    ref = pcm_dev(ctx, info["n_pcm_frames"]).clone()
    print(json.dumps({"workload": wl, "cfg": "sequential", "ms": round(ms0, 4), "stage_ms": [round(x, 4) for x in st0[:3]], "frames": info["n_frames"]}), flush=True)
    if QUICK:
        ctx.close(); del ref; torch.cuda.empty_cache(); continue
    for fpc in (24, 40, 48, 64):
        ctx.set_frames_per_cta(fpc)
        info, ms, st = measure(ctx, stream)
        eq = bool(torch.equal(pcm_dev(ctx, info["n_pcm_frames"]), ref))
        print(json.dumps({"workload": wl, "cfg": "sequential fpc=%d" % fpc, "ms": round(ms, 4), "stage_ms": [round(x, 4) for x in st[:3]], "pcm_equal": eq}), flush=True)
    ctx.set_frames_per_cta(32)
    sweep = [(c, p, 0, 0) for p in (0, 1) for c in (W2 // 2, W2, W3, 2 * W2, 2 * W3, 4 * W3)]
    sweep += [(c, p, 24 * 1024, 0) for p in (0, 1) for c in (W2, 2 * W2, 4 * W2)]                 # synthesis capped at 2 CTAs per SM
    sweep += [(c, 1, 0, 30 * 1024) for c in (W3, 2 * W3)] + [(c, 1, 24 * 1024, 30 * 1024) for c in (W2, 2 * W2)]   # K1 capped at 2 CTAs per SM
    best = None
    for chunk, prio, spad, kpad in sweep:
        ctx.set_overlap(chunk, prio, spad, kpad)
        info, ms, st = measure(ctx, stream)
        eq = bool(torch.equal(pcm_dev(ctx, info["n_pcm_frames"]), ref))
        print(json.dumps({"workload": wl, "cfg": "overlap chunk=%d prio=%d synth_pad=%d k1_pad=%d" % (chunk, prio, spad, kpad), "ms": round(ms, 4),
                          "vs_sequential": round(ms / ms0, 4), "pcm_equal": eq}), flush=True)
        if eq and (best is None or ms < best[0]): best = (ms, chunk, prio, spad, kpad)
    print(json.dumps({"workload": wl, "best_overlap": best, "sequential_ms": ms0}), flush=True)
    ctx.set_overlap(0)
    ctx.close(); del ref
    torch.cuda.empty_cache()
