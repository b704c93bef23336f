#!/usr/bin/env python3
"""bench.py -- headline benchmark of the MP3 Layer III granule decode path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl reference]

A step = one pass of the hot path (Huffman -> requantize/reorder/stereo/antialias -> IMDCT ->
polyphase -> int16 PCM) over one batch of synthetic frames per GPU.  Workload (BASELINE.json
configs[2], and per GPU of configs[4]): 1 000 000 frames of 44.1 kHz 320 kbps CBR joint stereo (MS),
bit reservoir in use; the stream is a 15 625-frame seeded block tiled 64x (distinct addresses, so
nothing is reused from L2; input 1.04 GB + output 4.6 GB per step >> the 126 MB L2).
Metric: decoded PCM sample-frames per second (1152 per MP3 frame), whole job over all N GPUs.

  value     kernels only, inputs resident in HBM, CUDA events, max over ranks
  e2e       the same work through the drop-in pdmp3_* C API with HOST buffers: pdmp3_feed() of the
            byte stream + pdmp3_read() into a host PCM buffer (parse, H2D, kernels, D2H all inside)
  roofline  dominant kernel: algorithmic bytes (frame bytes + 4608 PCM bytes per frame)
            per launch / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the UNMODIFIED reference (oracle/_ref) timed on this box's host cores
--impl reference: times only the reference CPU decoder (all host cores, forked processes).
"""
import argparse, ctypes as C, json, os, subprocess, sys, tempfile, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))       # tools/p3synth.py: the stream generator (never the test harness / oracle in the GPU arm)
import numpy as np

BLOCK = 15625
CFG = dict(bitrate_index=14, mode=1, mode_ext=2, blocks=0)       # configs[2]: 320 kbps CBR joint stereo (MS), long blocks
WORKLOAD = "1M-frame 44.1kHz 320kbps CBR joint-stereo(MS) synthetic stream, full on-device pipeline (BASELINE configs[2]; per GPU of configs[4])"
# --workload vbr: BASELINE configs[3], the divergence stress (not the headline line): VBR 32-320 kbps, long/short/mixed blocks, MS + intensity
CFG_VBR = dict(bitrate_index=0, mode=1, mode_ext=-1, blocks=1, overrun_pm=30)
WORKLOAD_VBR = "1M-frame 44.1kHz VBR 32-320kbps joint-stereo synthetic stream, mixed long/short/mixed blocks, MS+intensity, full on-device pipeline (BASELINE configs[3])"


def make_stream(n_frames, seed=1):
    import p3synth
    nblk = (n_frames + BLOCK - 1) // BLOCK
    blk, _ = p3synth.synth(min(BLOCK, n_frames), seed=seed, **CFG)
    if nblk == 1:
        return blk
    return np.tile(blk, nblk)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []; self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._rd, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _rd(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15); self.p.terminate()
        try: self.p.wait(timeout=2)
        except Exception: pass
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def _harness():
    """the test harness (oracle / compiled-reference bindings): ONLY the CPU-baseline legs import it"""
    t = os.path.join(ROOT, "tests")
    if t not in sys.path: sys.path.insert(0, t)
    import p3harness
    return p3harness


def _port_worker(args):
    H = _harness()
    seed, n = args
    s, _ = H.synth(n + 2, seed=seed, **CFG)
    t0 = time.perf_counter(); o = H.oracle_decode(s, lookahead=1152, taps=False); dt = time.perf_counter() - t0
    return o["n_frames"], dt


def ref_cpu_throughput(stream_block_frames, nprocs, frames_per_proc, exe_name="ref_bench"):
    """CPU arm: the unmodified reference (oracle/_ref/ref_bench, kind 'reference') on `nprocs` forked processes;
    if it did not travel with the repo, the oracle restatement (kind 'port') in a process pool.
    -> (sample-frames/s, frames, seconds, kind)"""
    import p3synth
    exe = os.path.join(ROOT, "oracle", "_ref", exe_name)
    if not os.path.exists(exe):
        if exe_name != "ref_bench":
            return None
        import multiprocessing as mp
        with mp.get_context("fork").Pool(nprocs) as pool:
            t0 = time.perf_counter(); res = pool.map(_port_worker, [(3, min(frames_per_proc, 1024))] * nprocs); wall = time.perf_counter() - t0
        frames = sum(r[0] for r in res); secs = max(r[1] for r in res)
        return frames * 1152 / secs, frames, wall, "port"
    s, _ = p3synth.synth(frames_per_proc + 2, seed=3, **CFG)
    d = tempfile.mkdtemp(prefix="p3bench", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    sp, op = os.path.join(d, "s.mp3"), os.path.join(d, "off.bin")
    big = np.tile(s, nprocs); big.tofile(sp)
    np.arange(nprocs + 1, dtype=np.uint64).__mul__(len(s)).astype("<u8").tofile(op)
    best = None
    for _ in range(2):
        out = subprocess.run([exe, sp, str(nprocs), op], stdout=subprocess.PIPE, text=True).stdout.split()
        frames, secs = int(out[0]), float(out[1])
        if best is None or secs < best[1]:
            best = (frames, secs)
    for f in (sp, op):
        os.unlink(f)
    os.rmdir(d)
    return best[0] * 1152 / best[1], best[0], best[1], "reference"


def bench_xr(a, rank, world, local):
    """--workload xr = BASELINE configs[1]: 65 536 frames of 44.1 kHz 128 kbps stereo, Huffman .. antialias done beforehand, the
    device runs the two transform kernels only (k_imdct + k_polyphase, P3_MODE_EXACT: bit-identical PCM).  `value`: spectra
    [65536][2][2][576] fp32 resident in HBM; e2e: p3_synth_from_xr() with the spectra in pinned host memory (604 MB H2D + 302 MB D2H)."""
    import torch, p3synth, pdmp3_b200
    torch.cuda.set_device(local)
    sys.stdout.flush(); _stdout_fd = os.dup(1); os.dup2(2, 1)
    nf = min(a.frames, 65536)
    blk, _ = p3synth.synth(min(4096, nf), seed=2, **p3synth.CONFIGS["cfg1_128k_stereo_long"])
    stream = np.tile(blk, (nf + 4095) // 4096) if nf > 4096 else blk
    ctx = pdmp3_b200.Context(local, pdmp3_b200.MODE_EXACT)
    parsed = pdmp3_b200.parse_stream(stream, lookahead=0)
    n_frames = parsed.n_pcm_frames
    ctx.upload(parsed); ctx.run(); ctx.sync()                # fills the device-resident spectra (the xr tap buffer of EXACT mode)
    ctx.time_xr(max(a.warmup, 3))
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    ms, st = ctx.time_xr(a.steps)
    torch.cuda.synchronize()
    e2e = None
    if not a.no_e2e:
        ctx.reset()
        pcm, taps = ctx.decode_parsed(parsed, taps=True)
        hx = torch.empty(taps["xr"].shape, dtype=torch.float32).pin_memory(); hx.numpy()[:] = taps["xr"]; del taps
        hp = torch.empty(n_frames * 1152 * 2, dtype=torch.int16).pin_memory()
        L = pdmp3_b200.lib(); times = []
        for it in range(1 + a.steps):
            ctx.reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
            rc = L.p3_synth_from_xr(ctx.h, hx.data_ptr(), C.byref(parsed.c), hp.data_ptr()); assert rc == 0, rc
            dt = time.perf_counter() - t0
            if it >= 1: times.append(dt)
        assert np.array_equal(hp.numpy().reshape(pcm.shape), pcm)
        dt = float(np.median(times))
        e2e = {"value": n_frames * 1152 / dt, "unit": "sample-frames/s", "h2d_bytes_per_step": int(hx.numel() * 4 + n_frames * 96),
               "d2h_bytes_per_step": int(hp.numel() * 2), "ms_per_step": 1e3 * dt,
               "api": "p3_synth_from_xr() with pinned host spectra and PCM buffers; H2D, k_imdct, k_polyphase, D2H inside the timed region"}
    clocks = sampler.stop()
    peak, peak_src = peaks()
    alg = 13824.0 * n_frames
    names = ["k_imdct", "k_polyphase"]; dom = int(np.argmax(st))
    roof = {"bound": "fp32 FFMA/LSU (direct-form transforms in the reference's summation order); labelled against HBM as SURVEY 8d asks",
            "kernel": names[dom], "achieved": alg / (st[dom] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (st[dom] * 1e-3) / 1e9 / peak,
            "traffic": None, "algorithmic_bytes": alg, "algorithmic_bytes_per_frame": 13824.0, "peak_source": peak_src,
            "whole_path": {"achieved": alg / (ms * 1e-3) / 1e9, "frac": alg / (ms * 1e-3) / 1e9 / peak}, "stage_ms": dict(zip(names, st))}
    sys.stdout.flush(); os.dup2(_stdout_fd, 1)
    v = n_frames * 1152 / (ms * 1e-3)
    print(json.dumps({"metric": "decoded_pcm_sample_frames_per_sec", "value": v, "unit": "sample-frames/s", "x_realtime_44k1": v / 44100.0,
                      "n_gpus": 1, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "65536-frame 44.1kHz 128kbps stereo batch, IMDCT+polyphase kernels only, spectra after antialias resident in HBM (BASELINE configs[1])",
                                 "frames_per_gpu": int(n_frames), "mode": "exact", "l2": "604 MB in + 302 MB out per step exceed the 126 MB L2; no flush needed"},
                      "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * a.steps, "roofline": roof, "cpu_baseline": None}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=1000000, help="frames per GPU per step")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mode", default="fast", choices=["exact", "fast"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--chunk", type=int, default=0, help="N > 1: frames per launch sequence and per PCM block on the wire (0: the library default, four waves of the synthesis kernel)")
    ap.add_argument("--workload", default="cbr320", choices=["cbr320", "vbr", "xr"], help="cbr320 = BASELINE configs[2] (headline); vbr = configs[3]; xr = configs[1] (transform kernels only)")
    a = ap.parse_args()
    global CFG, WORKLOAD
    if a.workload == "vbr": CFG, WORKLOAD = CFG_VBR, WORKLOAD_VBR
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    ncores = os.cpu_count() or 1

    if a.impl == "reference":
        if rank != 0:
            return 0
        per = 4096
        vals = []
        for i in range(max(a.warmup, 0) + max(a.steps, 1)):
            r = ref_cpu_throughput(BLOCK, ncores, per)
            if i >= a.warmup:
                vals.append(r)
        v = float(np.median([x[0] for x in vals])); fr = vals[0][1]
        # the other CPU baselines SURVEY 8d asks for, once each (bounded samples): one core, and the reference Makefile's own flags (-Os -ffast-math ...)
        variants = {}
        for key, exe, npr in (("O2_ieee_1core", "ref_bench", 1), ("stock_makefile_flags_allcores", "ref_bench_stock", ncores), ("stock_makefile_flags_1core", "ref_bench_stock", 1)):
            r = ref_cpu_throughput(BLOCK, npr, per, exe)
            if r: variants[key] = {"value": r[0], "cores": npr, "frames": r[1]}
        sample = "%d forked processes x %d frames of the same 320 kbps joint-stereo stream, pdmp3_read() 16 KiB / pdmp3_feed() 4096 B loop (pdmp3.c:2564-2584), -O2 IEEE build" % (ncores, per)
        print(json.dumps({"impl": "reference", "metric": "decoded_pcm_sample_frames_per_sec", "value": v, "unit": "sample-frames/s",
                          "x_realtime_44k1": v / 44100.0, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": 1e3 * fr * 1152 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "frames_per_step": fr},
                          "cpu_baseline": {"value": v, "unit": "sample-frames/s", "cores": ncores, "kind": vals[0][3], "sample": sample, "variants": variants},
                          "e2e": {"value": v, "unit": "sample-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return 0

    import torch
    import torch.distributed as dist
    import pdmp3_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the decoder has no CPU fallback")
    if a.workload == "xr":
        return bench_xr(a, rank, world, local) if rank == 0 else 0
    torch.cuda.set_device(local)
    # stdout carries exactly ONE JSON line: whatever libraries print there meanwhile (NCCL's version banner ...) goes to stderr
    sys.stdout.flush(); _stdout_fd = os.dup(1); os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nf = a.frames
    stream = make_stream(nf)
    # rank r > 0 decodes a shard of one long stream: 2 warm-up frames in front (reservoir bytes + filter state)
    warm = 2 if rank > 0 else 0
    if warm:
        blk_tail = make_stream(min(BLOCK, nf))
        fr_b = pdmp3_b200.parse_stream(blk_tail, lookahead=0).frames()      # product parser (not the oracle's copy)
        cut = int(fr_b["main_off"][-2]) - 36
        stream = np.concatenate([blk_tail[cut:], stream])
    ctx = pdmp3_b200.Context(local, pdmp3_b200.MODE_FAST if a.mode == "fast" else pdmp3_b200.MODE_EXACT)
    # nothing is parsed on the host: the bytes go to the device and the frame hop runs there (p3_hop.cu), as in pdmp3_read()
    info = ctx.upload_raw(stream, lookahead=0, warmup=warm); ctx.sync()
    info = ctx.upload_raw(stream, lookahead=0, warmup=warm); ctx.sync()         # (the second staging finds every buffer allocated: its hop time is the steady-state one)
    n_frames = info["n_pcm_frames"]; hop_ms = info["hop_ms"]
    class _P: pass
    parsed = _P(); parsed.n_frames = info["n_frames"]; parsed.n_pcm_frames = n_frames
    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
    for _ in range(max(a.warmup, 3)):
        ctx.run()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    ms_tot, ms_stage = ctx.time(a.steps)                    # CUDA events on the context stream, K steps
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = ctx.launch_count()
    t = torch.tensor([ms_tot], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n_frames * 1152 / (ms * 1e-3)

    # ---- N > 1: BASELINE configs[4] as ONE path -- the whole stream (world x frames) resident in rank 0's HBM, device hop + shard plan
    #      on rank 0, NCCL scatter of the byte ranges, every rank decodes its shard chunk by chunk, PCM blocks gathered into rank 0's HBM
    #      while later chunks decode (p3_sharded_decode, NCCL called from the C side).  This is `value` for N > 1; the number above
    #      (N independent replicas, no communication) stays as an extra key. ----
    sharded = None
    if world > 1:
        replicas = {"value": value, "ms_per_step": ms, "note": "N independent replicas of the single-GPU kernel sequence, no scatter / gather"}
        ids = [pdmp3_b200.dist_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sctx = pdmp3_b200.Context(local, pdmp3_b200.MODE_FAST if a.mode == "fast" else pdmp3_b200.MODE_EXACT)
        dd = pdmp3_b200.Dist(sctx, ids[0], rank, world)
        total = world * nf
        dev = None
        if rank == 0:
            big = make_stream(total)
            dev = torch.zeros(len(big) + 64, dtype=torch.uint8, device="cuda"); dev[:len(big)] = torch.from_numpy(big).cuda()
            big_bytes = len(big); del big
            torch.cuda.synchronize()
        ms_steps, res = [], None
        for it in range(max(a.warmup, 3) + a.steps):
            barrier()
            res = dd.sharded_decode(device_ptr=dev.data_ptr(), nbytes=big_bytes, chunk_frames=a.chunk) if rank == 0 else dd.sharded_decode(chunk_frames=a.chunk)
            if it >= max(a.warmup, 3): ms_steps.append(res["ms"])
        barrier()
        tsh = torch.tensor([float(np.mean(ms_steps))], dtype=torch.float64, device="cuda"); dist.all_reduce(tsh, op=dist.ReduceOp.MAX)
        ms_sh = float(tsh.item())
        tsc = torch.tensor([res["ms_scatter"]], dtype=torch.float64, device="cuda"); dist.all_reduce(tsc, op=dist.ReduceOp.MAX)
        # what landed on rank 0: the stream is a 15 625-frame block tiled, so from the second tile on every tile's PCM is the same bytes whichever rank decoded it
        tiles_ok = None
        if rank == 0:
            nb = 0; ptr = pdmp3_b200.lib().p3_batch_pcm_device(sctx.h, None)
            class _W: pass
            w = _W(); w.__cuda_array_interface__ = {"shape": (total * 4608,), "typestr": "|u1", "data": (ptr, False), "version": 2}
            pcm_all = torch.as_tensor(w, device="cuda")
            tb = BLOCK * 4608; ntile = total // BLOCK
            ref_tile = pcm_all[tb:2 * tb]
            tiles_ok = bool(all(torch.equal(pcm_all[k * tb:(k + 1) * tb], ref_tile) for k in range(2, ntile))) if ntile > 2 else None
            assert res["n_frames_total"] == total and tiles_ok is not False, "gathered PCM is wrong"
        # the floor: the same PCM bytes entering rank 0 with nothing else going on
        barrier()
        floor_ms = dd.measure_ingest(nf * 4608, 3)
        tfl = torch.tensor([floor_ms], dtype=torch.float64, device="cuda"); dist.all_reduce(tfl, op=dist.ReduceOp.MAX)
        floor_ms = float(tfl.item())
        tm = torch.tensor([res["ms_staged"], res["ms_decoded"], res["ms"]], dtype=torch.float64, device="cuda")
        tms = [torch.zeros_like(tm) for _ in range(world)]; dist.all_gather(tms, tm)
        sharded = {"ms_per_step": ms_sh, "gather_transport": dd.gather_transport(), "scatter_transport": "cuda-ipc copy-engine push out of rank 0, rank order" if (res["pad_"] & 2) else "nccl send/recv, rank order", "per_rank_ms_staged_decoded_total": [[round(float(x), 3) for x in t.tolist()] for t in tms], "frames_total": total, "chunk_frames": a.chunk, "scatter_ms": float(tsc.item()),
                   "bytes_to_rank0": nf * 4608 * (world - 1), "bytes_from_rank0": (big_bytes * (world - 1)) // world if rank == 0 else None,
                   "ingest_floor_ms": floor_ms, "ingest_floor_GBps": nf * 4608 * (world - 1) / floor_ms / 1e6, "time_over_floor": ms_sh / floor_ms,
                   "tiles_identical_on_rank0": tiles_ok, "launches_per_step_rank0": res["launches"],
                   "floor": "a bare grouped ncclRecv of the same PCM bytes from all peers into rank 0 (p3_dist_measure_ingest); B200_PROFILING.md quotes 770 GB/s for a peer copy"}
        value = total * 1152 / (ms_sh * 1e-3); ms = ms_sh; launches = res["launches"]
        dd.close(); sctx.close(); del dev
        sharded["replicas_no_collective"] = replicas

    # ---- end to end through the drop-in C API, host buffers ----
    e2e = None
    if not a.no_e2e:
        L = pdmp3_b200.lib()
        raw_bytes = len(stream)
        hin = torch.empty(raw_bytes, dtype=torch.uint8).pin_memory(); hin.numpy()[:] = stream
        out_bytes = (parsed.n_frames) * 4608
        hout = torch.empty(out_bytes, dtype=torch.uint8).pin_memory()
        # the decoder options a caller with a page-locked buffer uses: feed=borrow -- pdmp3_feed keeps the caller's pointer instead of
        # copying 1 GB into the handle's ring first (the copying feed, the reference's semantics, is timed next to it)
        runs = {}
        for label, extra in (("borrow", ",feed=borrow"), ("copy", "")):
            dec = pdmp3_b200.Decoder("b200:ring=%d,device=%d,mode=%s%s" % (raw_bytes + 4096 if not extra else 65536, local, a.mode, extra))
            times = []
            for it in range(2 + a.steps):
                dec.open_feed()
                barrier(); t0 = time.perf_counter()
                rc = L.pdmp3_feed(dec.h, hin.data_ptr(), raw_bytes); assert rc == 0, rc
                done = C.c_size_t(0)
                rc = L.pdmp3_read(dec.h, hout.data_ptr(), out_bytes, C.byref(done))
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
                if it >= 2: times.append((dt, done.value))
            dec.close()
            runs[label] = times
        times = runs["borrow"]
        dt_copy = float(np.median([x[0] for x in runs["copy"]]))
        assert runs["copy"][0][1] == runs["borrow"][0][1]
        dt = float(np.median([x[0] for x in times])); done_b = times[0][1]
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * (done_b / 4) / float(tt.item()), "unit": "sample-frames/s", "h2d_bytes_per_step": int(raw_bytes + raw_bytes // 32 + 8192 * ((parsed.n_frames + 32767) // 32768)),
               "d2h_bytes_per_step": int(done_b), "ms_per_step": 1e3 * float(tt.item()), "ms_per_step_copying_feed_this_rank": 1e3 * dt_copy,
               "api": "pdmp3_new(\"b200:feed=borrow,..\") + pdmp3_feed() + pdmp3_read() with pinned host buffers (feed=borrow: the handle decodes out of the caller's buffer; with the copying feed of the reference's semantics, \"b200:ring=<1 GB>\", see ms_per_step_copying_feed_this_rank); H2D of byte windows (each ~3 % larger than what it turns out to hold), frame hop + side info on the device, kernels, D2H inside the timed region; N > 1: every rank feeds and reads its own host buffers (a gather to rank 0 would only add rank 0's PCIe link as the limit)"}

    if rank != 0:
        if world > 1: dist.destroy_process_group()
        return 0
    peak, peak_src = peaks()
    frame_bytes = len(stream) / parsed.n_frames
    alg_bytes = (frame_bytes + 4608.0) * n_frames
    # stereo workloads: the packed-FFMA2 warp kernels, one per content class, all launched over the same grid (the CTAs of the other classes
    # return at once): the long-block cbr320 stream is decoded by k_synth_warp_lean, the mixed-block vbr stream by k_synth_warp_same
    synth = "k_synth_fast" if os.environ.get("P3_SYNTH") == "cta" else ("k_synth_warp_lean" if a.workload == "cbr320" else "k_synth_warp_same")
    names = ["k_compact", "k_huffman", synth, "-", "-"] if a.mode == "fast" else ["k_compact", "k_huffman", "k_requant", "k_imdct", "k_polyphase"]
    staged = sum(ms_stage) > 0                              # per-kernel times exist when the batch ran as one launch sequence
    dom = int(np.argmax(ms_stage)) if staged else 0
    dom_ms = ms_stage[dom] if staged else ms
    if not staged: names = ["whole launch sequence (chunked)"] + ["-"] * 4
    traffic = None
    tj = os.path.join(ROOT, "profiles", "traffic.json")       # dram__bytes_read+write per frame of each kernel, from the committed ncu captures
    if os.path.exists(tj):
        tr = json.load(open(tj)).get(names[dom])
        if tr: traffic = tr["dram_bytes_per_frame"] * n_frames
    roof = {"bound": "hbm (SURVEY 8d's label for the path; ncu shows the kernels bound by instruction issue and the FMA pipe, DRAM 17-19 % busy)", "kernel": names[dom], "achieved": alg_bytes / (dom_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": alg_bytes / (dom_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
            "algorithmic_bytes_per_frame": frame_bytes + 4608.0,
            "whole_path": {"achieved": world * alg_bytes / (ms * 1e-3) / 1e9, "frac": world * alg_bytes / (ms * 1e-3) / 1e9 / (world * peak),
                           "note": None if world == 1 else "N > 1: bound by rank 0's NVLink ingress (4608 B per frame from N-1 peers), see sharded.time_over_floor"},
            "stage_ms": {k: v for k, v in zip(names, ms_stage) if k != "-"}}
    cpu = None
    if not a.no_cpu and world == 1:
        r = ref_cpu_throughput(BLOCK, ncores, 4096)
        if r:
            cpu = {"value": r[0], "unit": "sample-frames/s", "cores": ncores, "kind": r[3],
                   "sample": "%d forked processes x 4096 frames of the same stream type, reference API loop, %.2f s wall" % (ncores, r[2])}
    sys.stdout.flush(); os.dup2(_stdout_fd, 1)
    print(json.dumps({"metric": "decoded_pcm_sample_frames_per_sec", "value": value, "unit": "sample-frames/s",
                      "x_realtime_44k1": value / 44100.0, "int16_samples_per_sec": 2 * value, "frames_per_sec": value / 1152,
                      "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                      "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "frames_per_gpu": int(n_frames), "mode": a.mode,
                                 "l2": "inputs+outputs (5.6 GB per step) exceed the 126 MB L2; no flush needed",
                                 "parallelism": ("1 GPU" if world == 1 else "frame-sharded x%d: NCCL ncclSend/ncclRecv scatter of compressed byte ranges from rank 0, chunked ncclSend/ncclRecv gather of PCM to rank 0 (overlapped with the decode of later chunks), no collective inside the decode" % world),
                                 "frames_total": int(world * n_frames), "host_parse_s": 0.0,
                                 "device_hop_ms": hop_ms, "device_hop": "frame hop of the whole stream on the device (k_hop_*), %d resolution round(s); outside the timed kernel sequence, inside e2e" % info["rounds"]},
                      "clocks": clocks, "e2e": e2e, "gpu_launches": launches * a.steps, "roofline": roof, "cpu_baseline": cpu,
                      "sharded": sharded}))
    sys.stdout.flush()
    if world > 1: dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
