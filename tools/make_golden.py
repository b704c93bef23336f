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
    g_hot=(20, dict(gain=208, peak_pm=600, blocks=1, mode=1, mode_ext=-1, seed=107)),     # near full scale, 7 % of the samples clipped
)
ONLY = sys.argv[1:]                                       # names to (re)generate; default: all
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
for name, (n, kw) in CASES.items():
    if ONLY and name not in ONLY: continue
    s, _ = H.synth(n, **kw)
    r = H.ref_decode(s)
    nf = r["n_frames"]
    nch = int(r["hdr"][0, 7])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), stream=s, n_frames=nf,
                        pcm=r["pcm"][:, :, :nch], is_huff=r["is_huff"], count1=r["count1"],
                        xr_ali_bits=r["xr_ali"].view(np.uint32), y_hyb_bits=r["y_hyb"].view(np.uint32), side=r["side"], hdr=r["hdr"])
    print(name, nf, "frames", len(s), "bytes")

# integer stages only (tests/golden/int_only/): streams outside the envelope in which the reference's float stages are defined
if not ONLY or "gi_empty" in ONLY:
    os.makedirs(os.path.join(ROOT, "tests", "golden", "int_only"), exist_ok=True)
    s, _ = H.synth(60, seed=108, **H.CONFIGS["cfg4_vbr_mixed"])
    t = H.empty_some_parts(s, every=5)                       # part2_3_length := 0 with scalefac_compress untouched
    r = H.ref_decode(t)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "int_only", "gi_empty.npz"), stream=t, n_frames=r["n_frames"],
                        is_huff=r["is_huff"], count1=r["count1"], scf_l=r["scf_l"], scf_s=r["scf_s"])
    print("gi_empty", r["n_frames"], "frames", len(t), "bytes")
