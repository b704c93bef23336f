"""GPU parity tests proper: the CUDA path (through the C-ABI) against the oracle restatement, the
compiled reference (oracle/_ref, when it travelled with the repo) and the committed fixtures.
Bit-exact for integer stages; P3_MODE_EXACT is also bit-exact in PCM; P3_MODE_FAST within 1 LSB."""
import glob, os
import numpy as np, pytest
import p3harness as H

pytestmark = pytest.mark.gpu

VARIANTS = dict(
    cfg1=dict(H.CONFIGS["cfg1_128k_stereo_long"]),
    cfg3=dict(H.CONFIGS["cfg3_320k_js_ms"]),
    cfg4=dict(H.CONFIGS["cfg4_vbr_mixed"]),
    mono=dict(mode=3, blocks=1, bitrate_index=7),
    k48=dict(sfreq=1, mode=1, mode_ext=-1, blocks=1, bitrate_index=11),
    k32=dict(sfreq=2, mode=1, mode_ext=-1, blocks=1, bitrate_index=12),
    crc=dict(crc=1, mode=1, mode_ext=3, blocks=1),
    c1b=dict(count1_b_pm=500, mode=1, mode_ext=-1, blocks=1),
    garbage=dict(garbage_pm=200, blocks=1),
    nores=dict(reservoir=0, blocks=1, bitrate_index=5),
    dual=dict(mode=2, blocks=1, overrun_pm=200),
    loud=dict(gain=200, blocks=1),
    lowrate=dict(bitrate_index=1, blocks=1, mode=1, mode_ext=-1),
    # near full scale (table precision "matters for loud signals"): PCM rms 15.8k / 21k LSB, 7 % / 23 % of the samples clipped
    hot=dict(gain=208, peak_pm=600, blocks=1, mode=1, mode_ext=-1),
    fullscale=dict(gain=215, peak_pm=1000, blocks=1),
)


def live_scalefactors(gc, nch):
    """masks [n,2,2,21] / [n,2,2,12,3] of the scalefactor cells Read_Main_L3 writes for each granule-channel
    (pdmp3.c:1382-1435): long blocks l[0..20]; short s[0..11]; mixed l[0..7] + s[3..11].  The other cells keep
    whatever an earlier frame left there in the reference (and in the oracle), while K1 writes zeros."""
    f = H.gc_fields(gc)
    short = (f[..., 4] == 1) & (f[..., 5] == 2)
    mixed = short & (f[..., 6] == 1)
    ml = np.zeros(gc.shape[:2] + (21,), bool); ms = np.zeros(gc.shape[:2] + (12, 3), bool)
    ml[~short] = True; ml[mixed, :8] = True
    ms[short & ~mixed] = True; ms[mixed, 3:] = True
    n = gc.shape[0]
    ml = ml.reshape(n, 2, 2, 21); ms = ms.reshape(n, 2, 2, 12, 3)
    ml[:, :, nch:] = False; ms[:, :, nch:] = False
    return ml, ms


def feq(a, b):
    """float arrays equal bit for bit, +0 == -0"""
    return ((a.view(np.uint32) == b.view(np.uint32)) | ((a == 0) & (b == 0)))


GOLD = sorted(glob.glob(os.path.join(H.ROOT, "tests", "golden", "*.npz")))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_exact_mode_matches_the_committed_golden_fixtures(gpu_ctx, path):
    """The CUDA path (P3_MODE_EXACT, through the C-ABI) against the committed outputs of the UNMODIFIED reference
    (tests/golden/*.npz: 7 stream types incl. mono, 48 kHz + CRC, count1 table B, near full scale): every stage bit for bit.
    Needs neither the oracle nor oracle/_ref at run time."""
    g = np.load(path)
    gpu_ctx.reset()
    pcm, t = gpu_ctx.decode(g["stream"], lookahead=1152, taps=True)
    H.check_against_golden(g, pcm, t["is_huff"], t["count1"], t["xr"], t["y"])


@pytest.mark.parametrize("name", list(VARIANTS))
def test_stage_taps_exact_vs_oracle(gpu_ctx, name):
    s, _ = H.synth(150, seed=21, **VARIANTS[name])
    o = H.oracle_decode(s, lookahead=1152)
    gpu_ctx.reset()
    pcm, t = gpu_ctx.decode(s, lookahead=1152, taps=True)
    n = o["n_frames"]
    assert pcm.shape[0] == n
    nch = pcm.shape[2]
    assert np.array_equal(t["is_huff"][:, :, :nch], o["is_huff"][:, :, :nch]), "Huffman output"
    assert np.array_equal(t["count1"][:, :, :nch], o["count1"][:, :, :nch]), "count1"
    ml, ms = live_scalefactors(o["gcs"], nch)
    assert np.array_equal(t["scf_l"][ml], o["scf_l"][ml]) and np.array_equal(t["scf_s"][ms], o["scf_s"][ms]), "scalefactors"
    if name in ("cfg3", "cfg4"): assert t["scf_l"][ml].any() and (name != "cfg4" or t["scf_s"][ms].any())     # (low bit rates: scalefac_compress 0)
    assert feq(t["xr"][:, :, :nch], o["xr_ali"][:, :, :nch]).all(), "requantize/reorder/stereo/antialias"
    assert feq(t["y"][:, :, :nch], o["y_hyb"][:, :, :nch]).all(), "hybrid synthesis"
    assert np.array_equal(pcm, o["pcm"]), "PCM"
    if H.have_ref():
        r = H.ref_decode(s, taps=False)
        ref = r["pcm"] if nch == 2 else r["pcm"][:, :, :1]
        assert np.array_equal(pcm, ref), "PCM vs compiled reference"


@pytest.mark.parametrize("name", ["cfg3", "cfg4", "crc"])
def test_empty_parts_integer_stages(gpu_ctx, name):
    """parts with part2_3_length == 0 but scalefac_compress != 0 (tests/test_cpu_oracle.py pins this rule to the compiled
    reference): scalefactor bits are still read, the next part starts behind them, count1 stays stale -- K1's
    scalefactors, spectra and count1 against the oracle, with the side info parsed on the host and on the device."""
    s, _ = H.synth(220, seed=33, **VARIANTS[name])
    t = H.empty_some_parts(s)
    o = H.oracle_decode(t, lookahead=1152)
    ml, ms = live_scalefactors(o["gcs"], 2)
    full = (H.gc_fields(o["gcs"])[..., 0] != 0).reshape(-1, 2, 2)     # K1 taps 0 for an empty part; the stale value (Q6) is looked up downstream (w3)
    for hop_only in (False, True):
        gpu_ctx.reset()
        pcm, tp = gpu_ctx.decode(t, lookahead=1152, taps=True, hop_only=hop_only)
        assert np.array_equal(tp["is_huff"], o["is_huff"]) and np.array_equal(tp["count1"][full], o["count1"][full]) and not tp["count1"][~full].any()
        assert np.array_equal(tp["scf_l"][ml], o["scf_l"][ml]) and np.array_equal(tp["scf_s"][ms], o["scf_s"][ms])


def test_exact_batches_with_carried_state(gpu_ctx):
    """Streaming continuity: four batches with state carried in the context == one batch, bit for bit
    (IMDCT overlap, polyphase history, bit reservoir tail, stale count1)."""
    import pdmp3_b200
    s, _ = H.synth(230, seed=12, **H.CONFIGS["cfg4_vbr_mixed"])
    gpu_ctx.reset(); a = gpu_ctx.decode(s, lookahead=0)
    gpu_ctx.reset()
    st = pdmp3_b200._binding.P3ParseState(0, 0, 0, -1, -1)
    pos, parts = 0, []
    while True:
        p = pdmp3_b200.parse_stream(s[pos:], lookahead=0, max_frames=61, state=st)
        if p.n_frames == 0: break
        for f in range(p.n_frames): p.c.frames[f].pcm_index = f
        parts.append(gpu_ctx.decode_parsed(p)); pos += p.consumed
    assert np.array_equal(a, np.concatenate(parts))
    o = H.oracle_decode(s, lookahead=0)
    assert np.array_equal(a, o["pcm"])


@pytest.mark.parametrize("mode", ["exact", "fast"])
def test_shards_with_halo_bit_identical(mode):
    """the concatenated PCM of an N-way frame-sharded decode (each shard primed by its warm-up
    frames) is bit-identical to the single-pass decode -- the multi-GPU correctness check, run here as
    three shards on one GPU with a fresh decoder state per shard."""
    import pdmp3_b200
    from pdmp3_b200 import shard
    ctx = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST if mode == "fast" else pdmp3_b200.MODE_EXACT)
    s, _ = H.synth(300, seed=23, **H.CONFIGS["cfg4_vbr_mixed"])
    whole = ctx.decode(s, lookahead=0)
    fr = pdmp3_b200.parse_stream(s, lookahead=0).frames()
    parts = []
    for p in shard.plan_shards(fr, 3):
        ctx.reset()
        parts.append(ctx.decode(s[p["byte_lo"]:p["byte_hi"]], lookahead=0, warmup=p["warmup"]))
        assert parts[-1].shape[0] == p["last"] - p["first"]
    assert np.array_equal(np.concatenate(parts), whole)
    ctx.close()


def test_full_size_properties():
    """BASELINE configs[1] size (65 536 frames) through size-independent properties: the stream is a
    4096-frame block tiled 16 times, so (a) encode->decode round trip: the Huffman stage returns exactly
    the encoded spectra for every tile, (b) PCM of tiles 2..16 is identical (same input, same carried
    state), (c) decoding twice gives the same bits (no races), (d) a tile decoded alone from zero state
    differs from a mid-stream tile only in its first frame (halo = 1 frame of filter state)."""
    import pdmp3_b200
    blk, iso = H.synth(4096, want_is=True, seed=77, **H.CONFIGS["cfg3_320k_js_ms"])
    s = np.tile(blk, 16)
    ctx = pdmp3_b200.Context(0, pdmp3_b200.MODE_EXACT)
    pcm, t = ctx.decode(s, lookahead=0, taps=True)
    assert pcm.shape[0] == 65536
    ih = t["is_huff"].reshape(16, 4096, 2, 2, 576)
    assert all(np.array_equal(ih[k], iso) for k in range(16))
    tiles = pcm.reshape(16, 4096, 1152, 2)
    assert all(np.array_equal(tiles[k], tiles[1]) for k in range(2, 16))
    ctx.reset(); again = ctx.decode(s, lookahead=0)
    assert np.array_equal(pcm, again)
    ctx.reset(); alone = ctx.decode(blk, lookahead=0)
    assert np.array_equal(alone[1:], tiles[1][1:]) and not np.array_equal(alone[0], tiles[1][0])
    ctx.close()


@pytest.mark.parametrize("name", ["cfg1", "cfg4", "mono", "k48"])
def test_config2_imdct_and_polyphase_only(gpu_ctx, name):
    """BASELINE configs[1]: Huffman .. antialias on the host (here: the oracle), IMDCT + polyphase on the device
    (p3_synth_from_xr): PCM bit-identical to the oracle's, first frame (zero state) included."""
    import pdmp3_b200
    s, _ = H.synth(400, seed=23, **VARIANTS[name])
    o = H.oracle_decode(s, lookahead=1152)
    p = pdmp3_b200.parse_stream(s, lookahead=1152)
    gpu_ctx.reset()
    pcm = gpu_ctx.synth_from_xr(o["xr_ali"], p)
    assert pcm.shape[0] == o["n_frames"]
    assert np.array_equal(pcm, o["pcm"][:, :, :pcm.shape[2]])
    assert gpu_ctx.launch_count() == 2


def test_config2_full_size_65536_frames(gpu_ctx):
    """configs[1] at its full size (65 536 frames of 128 kbps stereo = 262 144 granule-channels, 604 MB of fp32 spectra
    in, 302 MB of PCM out): the transform-only entry point fed with the spectra of the full pipeline gives the full
    pipeline's PCM, bit for bit, and the 4096-frame blocks the stream is tiled from decode identically."""
    import pdmp3_b200
    blk, _ = H.synth(4096, seed=2, **H.CONFIGS["cfg1_128k_stereo_long"])
    s = np.tile(blk, 16)
    p = pdmp3_b200.parse_stream(s, lookahead=0)
    assert p.n_frames == 65536
    gpu_ctx.reset(); pcm, t = gpu_ctx.decode_parsed(p, taps=True)
    xr = t["xr"]; del t
    gpu_ctx.reset(); pcm2 = gpu_ctx.synth_from_xr(xr, p)
    assert np.array_equal(pcm, pcm2)
    tiles = pcm2.reshape(16, 4096, 1152, 2)
    for k in range(2, 16):
        assert np.array_equal(tiles[k], tiles[1]), "tile %d" % k
