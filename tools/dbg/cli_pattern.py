"""throughput of the reference's own caller pattern (pdmp3_read 16 KiB / pdmp3_feed 4096 B, default 16 KiB ring) through
the drop-in library and through the compiled reference (1 core)"""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..")); sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
import p3harness as H, pdmp3_b200
from test_gpu_api import RefApi, cli_loop, REFLIB
s, _ = H.synth(4000, seed=5, **H.CONFIGS["cfg3_320k_js_ms"])
for name, mk in (("b200 (fast)", lambda: pdmp3_b200.Decoder()), ("b200 ring=1MiB", lambda: pdmp3_b200.Decoder("b200:ring=1048576")), ("reference, 1 core", (lambda: RefApi()) if os.path.exists(REFLIB) else None)):
    if mk is None: continue
    d = mk(); cli_loop(d, s[:200000])
    fs = 65536 if "1MiB" in name else 4096
    t0 = time.perf_counter(); pcm, tr = cli_loop(d, s, feedsize=fs); dt = time.perf_counter() - t0
    d.close()
    print("%-20s %6d frames in %.3f s = %8.0f frames/s (%d calls)" % (name, len(pcm) // 4608, dt, len(pcm) / 4608 / dt, len(tr)))
