import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import p3harness as H, pdmp3_b200
from test_gpu_parity import VARIANTS
ctx = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST)
for name in sys.argv[1:] or ["cfg3", "cfg1", "cfg4"]:
    s, _ = H.synth(300, seed=17, **VARIANTS[name])
    o = H.oracle_decode(s, lookahead=0)
    ctx.reset(); ctx.set_synth_kernel(1); ref = ctx.decode(s, lookahead=0)
    for fpw in (32, 5, 1):
        ctx.reset(); ctx.set_synth_kernel(0); ctx.set_frames_per_cta(fpw); got = ctx.decode(s, lookahead=0)
        d = got.astype(np.int32) - ref.astype(np.int32)
        nz = np.argwhere(d != 0)
        do = np.abs(got.astype(np.int32) - o["pcm"].astype(np.int32))
        print(name, "fpw", fpw, "differ", len(nz), "max|d|", np.abs(d).max(), "vs oracle max", do.max(), "eq frac", (do == 0).mean(),
              "old-vs-oracle max", np.abs(ref.astype(np.int32) - o["pcm"].astype(np.int32)).max())
        if len(nz):
            fr = np.unique(nz[:, 0]); print("  frames:", fr[:20], "n", len(fr)); print("  first:", nz[:8].tolist(), d[tuple(nz[0])])
            f0 = nz[0][0]; sl = (nz[nz[:, 0] == f0][:, 1]); print("  in frame", f0, "sample idx range", sl.min(), sl.max(), "channels", np.unique(nz[nz[:,0]==f0][:,2]))
