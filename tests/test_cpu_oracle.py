"""CPU tests (-m "not gpu"): the oracle restatement against (a) the committed golden fixtures produced by
the unmodified reference and (b) the compiled reference itself when oracle/_ref travelled with the repo;
the host parser; the generator round trip.  No GPU, no /root/reference at run time."""
import glob, os
import numpy as np, pytest
import p3harness as H

GOLD = sorted(glob.glob(os.path.join(H.ROOT, "tests", "golden", "*.npz")))


def feq_bits(a_bits, b):
    bb = b.view(np.uint32)
    return (a_bits == bb) | (((a_bits & 0x7fffffff) == 0) & ((bb & 0x7fffffff) == 0))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_golden(path):
    g = np.load(path)
    o = H.oracle_decode(g["stream"], lookahead=1152)
    n = int(g["n_frames"])
    assert o["n_frames"] == n
    nch = g["pcm"].shape[2]
    assert np.array_equal(o["is_huff"][:, :, :nch], g["is_huff"][:, :, :nch])
    assert np.array_equal(o["count1"][:, :, :nch], g["count1"][:, :, :nch])
    assert feq_bits(g["xr_ali_bits"][:, :, :nch], o["xr_ali"][:, :, :nch]).all()
    assert feq_bits(g["y_hyb_bits"][:, :, :nch], o["y_hyb"][:, :, :nch]).all()
    assert np.array_equal(o["pcm"], g["pcm"])
    # parser vs the reference's side-info parse
    side = H.gc_fields(o["gcs"]).reshape(n, 2, 2, 20)
    ref = g["side"]
    ws = ref[..., 4] == 1
    for k in range(18):
        m = side[..., k] != ref[..., k]
        if k in (6, 10, 11, 12): m &= ws          # only parsed when win_switch=1 (stale otherwise, pdmp3.c:1174-1180)
        if k == 9: m &= ~ws                       # table_select[2] only parsed when win_switch=0
        assert not m[:, :, :nch].any(), "side-info field %d" % k
    assert np.array_equal(o["frames"]["main_begin"], g["hdr"][:, 6])


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_golden_checker_on_the_oracle(path):
    """the fixture checker the GPU parity test uses (p3harness.check_against_golden), run here on the oracle's arrays"""
    g = np.load(path)
    o = H.oracle_decode(g["stream"], lookahead=1152)
    H.check_against_golden(g, o["pcm"], o["is_huff"], o["count1"], o["xr_ali"], o["y_hyb"])
    bad = o["pcm"].copy(); bad[0, 5, 0] ^= 1                   # ... and it does notice a single flipped PCM bit
    with pytest.raises(AssertionError):
        H.check_against_golden(g, bad, o["is_huff"], o["count1"], o["xr_ali"], o["y_hyb"])


def test_oracle_matches_golden_empty_parts():
    """integer stages of a stream with empty parts (part2_3_length == 0, scalefac_compress != 0) as the unmodified
    reference decodes them (tests/golden/int_only/gi_empty.npz, tools/make_golden.py)"""
    g = np.load(os.path.join(H.ROOT, "tests", "golden", "int_only", "gi_empty.npz"))
    o = H.oracle_decode(g["stream"], lookahead=1152)
    assert o["n_frames"] == int(g["n_frames"])
    f = H.gc_fields(o["gcs"])
    assert ((f[..., 0] == 0) & (f[..., 3] != 0)).sum() >= 5
    for k in ("scf_l", "scf_s", "is_huff", "count1"):
        assert np.array_equal(o[k], g[k]), k


SWEEP = dict(
    cfg1=dict(H.CONFIGS["cfg1_128k_stereo_long"]), cfg3=dict(H.CONFIGS["cfg3_320k_js_ms"]), cfg4=dict(H.CONFIGS["cfg4_vbr_mixed"]),
    mono=dict(mode=3, blocks=1, bitrate_index=7), k48=dict(sfreq=1, mode=1, mode_ext=-1, blocks=1, bitrate_index=11),
    k32=dict(sfreq=2, mode=1, mode_ext=-1, blocks=1, bitrate_index=12), crc=dict(crc=1, mode=1, mode_ext=3, blocks=1),
    c1b=dict(count1_b_pm=500, mode=1, mode_ext=-1, blocks=1), garbage=dict(garbage_pm=200, blocks=1),
    nores=dict(reservoir=0, blocks=1, bitrate_index=5), dual=dict(mode=2, blocks=1, overrun_pm=200),
    loud=dict(gain=200, blocks=1), lowrate=dict(bitrate_index=1, blocks=1, mode=1, mode_ext=-1),
    hot=dict(gain=208, peak_pm=600, blocks=1, mode=1, mode_ext=-1), fullscale=dict(gain=215, peak_pm=1000, blocks=1),
)


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("name", list(SWEEP))
def test_oracle_matches_compiled_reference(name):
    """Pins the restatement: every stage tap and the PCM, bit for bit, on a fresh seeded stream."""
    s, _ = H.synth(120, seed=77, **SWEEP[name])
    r = H.ref_decode(s); o = H.oracle_decode(s, lookahead=1152)
    assert r["n_frames"] == o["n_frames"] > 100
    for k in ("is_huff", "count1", "xr_req", "xr_reo", "xr_ste", "xr_ali", "y_hyb"):
        a, b = r[k], o[k]
        ok = feq_bits(a.view(np.uint32), b) if a.dtype == np.float32 else (a == b)
        assert ok.all(), k
    nch = o["pcm"].shape[2]
    assert np.array_equal(r["pcm"][:, :, :nch], o["pcm"])


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("name", ["cfg3", "cfg4", "crc"])
def test_empty_parts_follow_the_reference(name):
    """A part with part2_3_length == 0 (outside the generator's envelope, G5): the reference still reads its scalefactor
    bits (scalefac_compress is whatever the stream says) and does NOT reposition, so the following parts of the frame start
    behind those bits; count1 of the slot stays stale (Q6).  Integer stages -- scalefactors, Huffman output, count1 --
    must match the compiled reference bit for bit (the float stages are outside the envelope here: the shifted bits no
    longer keep the sfb-0 scalefactors at 0, which is what makes Q5's out-of-bounds reads defined)."""
    s, _ = H.synth(220, seed=33, **SWEEP[name])
    t = H.empty_some_parts(s)
    fr, gc, info = H.parse(t, lookahead=1152)
    f = H.gc_fields(gc)
    empties = (f[..., 0] == 0)
    assert empties.sum() > 20 and (f[..., 3][empties] != 0).any(), "need empty parts with scalefac_compress != 0"
    r = H.ref_decode(t); o = H.oracle_decode(t, lookahead=1152)
    assert r["n_frames"] == o["n_frames"] > 200
    for k in ("scf_l", "scf_s", "is_huff", "count1"):
        assert np.array_equal(r[k], o[k]), k


def test_generator_roundtrip():
    """encode -> decode: the oracle's Huffman stage returns exactly the spectra the generator encoded."""
    s, iso = H.synth(200, want_is=True, seed=5, **H.CONFIGS["cfg3_320k_js_ms"])
    o = H.oracle_decode(s, lookahead=0)
    n = o["n_frames"]
    assert n == 200
    assert np.array_equal(o["is_huff"], iso[:n])
    assert np.abs(iso).max() > 1000            # linbits escapes were exercised


def test_stream_coverage():
    """The VBR/mixed stream really contains what config 4 promises."""
    s, _ = H.synth(600, seed=9, **H.CONFIGS["cfg4_vbr_mixed"])
    fr, gc, info = H.parse(s, lookahead=0)
    f = H.gc_fields(gc)
    bt = f[..., 5][f[..., 4] == 1]
    assert set(np.unique(bt)) == {1, 2, 3}
    assert (f[..., 6] == 1).any()                                   # mixed blocks
    assert set(np.unique(fr["mode_ext"])) == {0, 1, 2, 3}          # MS and intensity stereo
    assert len(np.unique(fr["bitrate_kbps"])) == 14                 # VBR over all 14 bitrates
    assert fr["main_begin"].max() > 300 and (fr["scfsi"] != 0).any()
    assert (fr["flags"] & 6).sum() == 0                             # nothing flagged NODATA / BAD


def test_lookahead_rule_frame_counts():
    """Q7 (pdmp3.c:2445): a frame is read only when >= 1152 bytes are buffered: 10 frames in -> 8 / 9 out."""
    for kw, expect in ((dict(bitrate_index=9), 8), (dict(bitrate_index=14), 9)):
        s, _ = H.synth(10, seed=2, reservoir=0, **kw)
        fr, gc, info = H.parse(s, lookahead=1152)
        assert info["n_frames"] == expect
        fr, gc, info = H.parse(s, lookahead=0)
        assert info["n_frames"] == 10
        if H.have_ref():
            assert H.ref_decode(s, taps=False)["n_frames"] == expect


def test_parser_rejects_and_resyncs():
    s, _ = H.synth(30, seed=4, garbage_pm=500)
    fr, gc, info = H.parse(s, lookahead=0)
    assert info["n_frames"] == 30
    junk = np.full(5000, 0x55, np.uint8)
    fr, gc, info = H.parse(junk, lookahead=0)
    assert info["n_frames"] == 0 and info["stop"] == 2              # no header within a frame's length (pdmp3.c:1337)
    # truncated last frame is not parsed
    s2, _ = H.synth(5, seed=4)
    fr, gc, info = H.parse(s2[:-7], lookahead=0)
    assert info["n_frames"] == 4


def test_parser_threads_agree():
    s, _ = H.synth(6000, seed=8, **H.CONFIGS["cfg4_vbr_mixed"])
    a = H.parse(s, lookahead=0, nthreads=1); b = H.parse(s, lookahead=0, nthreads=8)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_bad_side_info_flagged():
    s, _ = H.synth(6, seed=4, reservoir=0)
    fr, gc, info = H.parse(s, lookahead=0)
    t = s.copy()
    off = int(fr["main_off"][2]) - 32
    # big_values of granule 0 channel 0 := 511 (> 288): bits 32..40 of the side info
    v = int.from_bytes(t[off + 2:off + 10].tobytes(), "big")
    v |= 0x1ff << (64 - 16 - 9 - 12 - 4 + 4 - 13 + 13)           # keep simple: force all nine big_values bits on
    t[off + 4] |= 0xff; t[off + 5] |= 0x80
    fr2, gc2, info2 = H.parse(t, lookahead=0)
    assert fr2["flags"][2] & 4
    o = H.oracle_decode(t, lookahead=0)
    assert o["n_frames"] == 6                                      # decodes (as silence for that frame) without crashing


def test_hop_only_parse_gives_the_same_frames():
    """p3_parse_opts.hop_only (side info left to the device parser): same frame descriptors except scfsi and the
    validity flag, which the device fills in; the granule descriptors stay zero on the host."""
    import pdmp3_b200
    s, _ = H.synth(300, seed=12, **H.CONFIGS["cfg4_vbr_mixed"])
    a = pdmp3_b200.parse_stream(s, lookahead=1152)
    b = pdmp3_b200.parse_stream(s, lookahead=1152, hop_only=True)
    assert a.n_frames == b.n_frames and a.consumed == b.consumed and a.n_pcm_frames == b.n_pcm_frames
    fa, fb = a.frames(), b.frames()
    for name in fa.dtype.names:
        if name in ("scfsi", "flags"): continue
        assert np.array_equal(fa[name], fb[name]), name
    assert np.array_equal(fa["flags"] & 0xfb, fb["flags"])          # P3_FRAME_BAD (4) comes from the side info
    assert b.c.hop_only == 1 and not b.gcs().any()
