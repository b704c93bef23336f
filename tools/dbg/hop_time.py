"""device hop at scale: python tools/dbg/hop_time.py [frames] [vbr]   (run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split)"""
import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import p3synth, pdmp3_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
cfg = dict(bitrate_index=0, mode=1, mode_ext=-1, blocks=1, overrun_pm=30) if "vbr" in sys.argv else dict(bitrate_index=14, mode=1, mode_ext=2, blocks=0)
blk, _ = p3synth.synth(15625, seed=1, **cfg)
s = np.tile(blk, (n + 15624) // 15625)
ctx = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST)
for it in range(3):
    info = ctx.upload_raw(s, lookahead=0); ctx.sync()
    print("frames %d bytes %d hop_ms %.3f rounds %d" % (info["n_frames"], len(s), info["hop_ms"], info["rounds"]), flush=True)
ctx.close()
