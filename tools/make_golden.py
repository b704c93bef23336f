#!/usr/bin/env python3
"""Generate tests/golden/*.npz: small seeded streams and what the UNMODIFIED reference
(oracle/_ref/libref_taps.so, built from /root/reference by oracle/Makefile) decodes them to.
Run in the build container only; the fixtures are committed so that the checks also run where
neither /root/reference nor oracle/_ref exists."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import p3harness as H

CASES = dict(
    g_cfg1=(26, dict(H.CONFIGS["cfg1_128k_stereo_long"], seed=101)),
    g_cfg3=(18, dict(H.CONFIGS["cfg3_320k_js_ms"], seed=102)),
    g_cfg4=(40, dict(H.CONFIGS["cfg4_vbr_mixed"], seed=103)),
    g_mono=(30, dict(mode=3, blocks=1, bitrate_index=7, seed=104)),
    g_48k_crc=(24, dict(sfreq=1, crc=1, mode=1, mode_ext=-1, blocks=1, bitrate_index=11, seed=105)),
    g_c1b=(24, dict(count1_b_pm=500, mode=1, mode_ext=-1, blocks=1, seed=106)),
)
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
for name, (n, kw) in CASES.items():
    s, _ = H.synth(n, **kw)
    r = H.ref_decode(s)
    nf = r["n_frames"]
    nch = int(r["hdr"][0, 7])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), stream=s, n_frames=nf,
                        pcm=r["pcm"][:, :, :nch], is_huff=r["is_huff"], count1=r["count1"],
                        xr_ali_bits=r["xr_ali"].view(np.uint32), y_hyb_bits=r["y_hyb"].view(np.uint32), side=r["side"], hdr=r["hdr"])
    print(name, nf, "frames", len(s), "bytes")
