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
