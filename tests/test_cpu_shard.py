"""Multi-rank host logic on CPU: world_size-2 gloo.  The sharded decode (scatter bytes, decode shard with
warm-up, gather PCM) must reproduce the single-rank PCM bit for bit.  The decode callable here is the CPU
oracle (this is a test of the sharding/halo logic; the GPU path plugs in pdmp3_b200.Context.decode)."""
import os, sys
import numpy as np, pytest
import p3harness as H


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, H.ROOT)
    from pdmp3_b200 import shard
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    s, _ = H.synth(96, seed=17, **H.CONFIGS["cfg4_vbr_mixed"])
    dec = lambda b, warm: H.oracle_decode(b, lookahead=0, taps=False, warmup=warm)["pcm"]
    out = shard.decode_sharded(s if rank == 0 else None, lambda x: H.parse(x, lookahead=0)[0], dec, rank, world)
    if rank == 0:
        whole = H.oracle_decode(s, lookahead=0, taps=False)["pcm"]
        q.put(bool(out.shape == whole.shape and np.array_equal(out, whole)))
    dist.destroy_process_group()


def test_two_rank_sharded_decode_bit_identical():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps: p.start()
    for p in ps: p.join(120)
    assert all(p.exitcode == 0 for p in ps)
    assert q.get(timeout=5) is True


def test_plan_covers_reservoir():
    sys.path.insert(0, H.ROOT)
    from pdmp3_b200 import shard
    s, _ = H.synth(400, seed=5, **H.CONFIGS["cfg4_vbr_mixed"])
    fr, gc, info = H.parse(s, lookahead=0)
    for world in (2, 3, 8):
        plans = shard.plan_shards(fr, world)
        assert plans[0]["first"] == 0 and plans[-1]["last"] == len(fr)
        whole = H.oracle_decode(s, lookahead=0, taps=False)["pcm"]
        for p in plans[1:]:
            assert p["warmup"] >= 1
            sub = s[p["byte_lo"]:p["byte_hi"]]
            o = H.oracle_decode(sub, lookahead=0, taps=False, warmup=p["warmup"])["pcm"]
            assert np.array_equal(o, whole[p["first"]:p["last"]]), (world, p)


def test_chunk_schedule_of_the_sharded_decode():
    """p3_sharded_decode cuts a shard into launch sequences (C/4, C/4, C/2, C ... C, C/2, C/4, C/4; uniform for short shards).
    Every rank derives the schedule from the plan alone, so it must be a partition of [0, n) into chunks of at most C frames that
    start on multiples of 32 (K1's frame groups), for any shard length."""
    import ctypes as C, pdmp3_b200
    L = pdmp3_b200.lib()
    L.p3_dist_chunk_start.restype = C.c_int64; L.p3_dist_chunk_start.argtypes = [C.c_int64] * 3
    import random
    rng = random.Random(5)
    cases = [(c, n) for c in (32, 64, 96, 256, 4096, 227328, 262144) for n in (1, 31, 32, 33, 700, 3 * c - 1, 3 * c, 3 * c + 1, 8 * c + 17, 1000002)]
    cases += [(32 * rng.randint(1, 9000), rng.randint(1, 3000000)) for _ in range(300)]
    for c, n in cases:
        prev, j, sizes = 0, 1, []
        while True:
            f = L.p3_dist_chunk_start(j, c, n)
            assert prev < f <= n, (c, n, j, prev, f)
            assert prev % 32 == 0 and f - prev <= c, (c, n, j, prev, f)
            sizes.append(f - prev); prev = f; j += 1
            if f == n: break
            assert j < 200000
        assert L.p3_dist_chunk_start(0, c, n) == 0 and L.p3_dist_chunk_start(j + 3, c, n) == n
        if n >= 3 * c and c >= 128: assert sizes[0] <= c // 4 and sizes[-1] <= c // 4 + 64, (c, n, sizes[:4], sizes[-4:])
