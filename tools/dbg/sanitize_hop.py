"""small device-hop + raw decode + world-of-one sharded decode runs for compute-sanitizer (memcheck / racecheck)"""
import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import p3harness as H, pdmp3_b200
from test_gpu_parity import VARIANTS
from test_gpu_hop import parallel_false_chain
ctx = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST)
for name in ("cfg4", "garbage", "mono"):
    s, _ = H.synth(160, seed=5, **VARIANTS[name])
    a = ctx.decode(s, lookahead=1152); ctx.reset()
    b, info = ctx.decode_raw(s, lookahead=1152); ctx.reset()
    assert np.array_equal(a, b), name
    for cut in (3, 500, len(s) - 9): ctx.upload_raw(s[:cut]); ctx.reset()
    print(name, info, flush=True)
s, _ = H.synth(200, seed=6, reservoir=0, **H.CONFIGS["cfg3_320k_js_ms"])
info = ctx.upload_raw(parallel_false_chain(s)); print("adversarial", info, flush=True)
d = pdmp3_b200.Dist(ctx, pdmp3_b200.dist_unique_id(), 0, 1)
s, _ = H.synth(300, seed=7, **H.CONFIGS["cfg4_vbr_mixed"])
res = d.sharded_decode(stream=s, chunk_frames=64); got = d.pcm(res)
one = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST); want = one.decode(s, lookahead=0)
assert np.array_equal(got, want)
print("sharded world-of-one ok", res["chunks"], flush=True)
d.close(); ctx.close(); one.close()
print("done")
