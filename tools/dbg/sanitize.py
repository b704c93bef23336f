"""small decodes of every stream type for compute-sanitizer (memcheck / racecheck / synccheck)"""
import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import p3harness as H, pdmp3_b200
from test_gpu_parity import VARIANTS
names = sys.argv[1:] or ["cfg3", "cfg4", "mono", "k48"]
NF = int(os.environ.get("P3_SAN_FRAMES", "70"))
for mode in (pdmp3_b200.MODE_FAST, pdmp3_b200.MODE_EXACT):
    ctx = pdmp3_b200.Context(0, mode)
    for n in names:
        s, _ = H.synth(NF, seed=3, **VARIANTS[n])
        a = ctx.decode(s, lookahead=0, hop_only=True)
        ctx.reset(); ctx.set_frames_per_cta(5); b = ctx.decode(s, lookahead=0); ctx.set_frames_per_cta(32); ctx.reset()
        assert np.array_equal(a, b), n
        print(mode, n, a.shape, flush=True)
    ctx.close()
print("done")
