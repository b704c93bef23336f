"""The frame hop on the device (p3_hop.cu; SURVEY 8f-1, Search_Header / Read_Header pdmp3.c:1252-1340, Read_Frame 1217-1244,
reservoir rule of Get_Main_Data 1101-1120): the p3_frame[] array it produces must equal, byte for byte, what the sequential host
hop of p3_parse.c produces -- on every stream type, with junk between frames, CRC words, truncated ends, format changes, frame
limits, adversarial false syncs -- and decoding from raw bytes must give the PCM of the host-parsed path."""
import numpy as np, pytest
import p3harness as H
from test_gpu_parity import VARIANTS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import pdmp3_b200
    c = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST)
    yield c
    c.close()


def compare_hop(ctx, s, lookahead=0, max_frames=0, warmup=0, state=None):
    """device hop (+ device side info) against the host parser; returns the device info dict"""
    import pdmp3_b200
    B = pdmp3_b200._binding
    st_h = B.P3ParseState(*state) if state else None
    st_d = B.P3ParseState(*state) if state else None
    host = pdmp3_b200.parse_stream(s, lookahead=lookahead, max_frames=max_frames, warmup=warmup, state=st_h)
    ctx.reset()
    info = ctx.upload_raw(s, lookahead=lookahead, max_frames=max_frames, warmup=warmup, state=st_d)
    assert info["n_frames"] == host.n_frames, (info, host.n_frames)
    assert info["n_pcm_frames"] == host.n_pcm_frames
    assert info["consumed"] == host.consumed, (info, host.consumed)
    assert info["stop"] == host.stop, (info, host.stop)
    n = host.n_frames
    if n:
        assert info["nch"] == host.nch
        ctx.run(); ctx.sync()                                  # k_sideinfo fills scfsi / the BAD flag / the granule descriptors
        fr_d, gc_d = ctx.download_desc(n)
        fr_h, gc_h = host.frames(), host.gcs()
        for name in fr_h.dtype.names:
            assert np.array_equal(fr_h[name], fr_d[name]), "frame field %s" % name
        live = [k for k in range(4) if (k & 1) < host.nch]
        assert np.array_equal(gc_h[:, live], gc_d[:, live]), "granule-channel descriptors"
    return info


@pytest.mark.parametrize("lookahead", [0, 1152])
@pytest.mark.parametrize("name", list(VARIANTS))
def test_device_hop_equals_host_hop(ctx, name, lookahead):
    s, _ = H.synth(400, seed=61, **VARIANTS[name])
    info = compare_hop(ctx, s, lookahead=lookahead)
    assert info["n_frames"] >= 380                                # (the 1152-byte look-ahead rule keeps the last frames back, more of them at low bit rates)


def test_limits_truncation_and_stop_codes(ctx):
    s, _ = H.synth(300, seed=62, **H.CONFIGS["cfg4_vbr_mixed"])
    for mf in (1, 2, 31, 32, 33, 299, 300, 301):
        info = compare_hop(ctx, s, max_frames=mf)
        assert info["stop"] == (1 if mf <= 300 else 0)
    for cut in (1, 3, 4, 5, 37, 500, len(s) - 7, len(s) - 1):   # truncated streams: incomplete last frame, fewer than 4 bytes, ...
        compare_hop(ctx, s[:cut]); compare_hop(ctx, s[:cut], lookahead=1152)
    assert compare_hop(ctx, s[:0])["n_frames"] == 0
    junk = np.full(5000, 0x55, np.uint8)
    info = compare_hop(ctx, junk)
    assert info["n_frames"] == 0 and info["stop"] == 2           # no header within a frame's length (pdmp3.c:1337)
    info = compare_hop(ctx, np.concatenate([s, junk]))
    assert info["n_frames"] == 300 and info["stop"] == 2
    compare_hop(ctx, np.concatenate([junk[:700], s]))            # leading junk shorter than the resync window
    compare_hop(ctx, s, warmup=2); compare_hop(ctx, s, warmup=5, max_frames=40)
    compare_hop(ctx, s[int(1e4):], state=(123456, 300, 17, 2, 0))    # carried parser state: main_pos base, reservoir level, pcm slots


def test_format_change_ends_the_batch(ctx):
    """stop = 3: the channel count / sample rate changes -> the batch ends in front of that header (consumed = its position)"""
    a, _ = H.synth(90, seed=63, **H.CONFIGS["cfg3_320k_js_ms"])
    b, _ = H.synth(90, seed=64, mode=3, bitrate_index=7)
    c, _ = H.synth(60, seed=65, sfreq=1, bitrate_index=11)
    for s in (np.concatenate([a, b]), np.concatenate([b, a, c]), np.concatenate([a, c, a])):
        info = compare_hop(ctx, s)
        assert info["stop"] == 3 and info["n_frames"] == 90
        info = compare_hop(ctx, s, max_frames=90)                 # the frame limit is tested first (p3_parse.c loop head)
        assert info["stop"] == 1
        rest = s[info["consumed"]:]
        compare_hop(ctx, rest)


def parallel_false_chain(s, d=600):
    """copy every frame's header to offset d inside the same frame: a second, self-consistent chain of valid headers with the
    same frame lengths that never meets the true one.  Speculative entries that find a false header first follow the false
    chain through their whole segment: the list-rewrite path and several resolution rounds."""
    fr, _, _ = H.parse(s, lookahead=0)
    t = s.copy()
    for f in fr:
        h = int(f["main_off"]) - 36
        t[h + d:h + d + 4] = s[h:h + 4]
    return t


def test_adversarial_false_sync_chain(ctx):
    s, _ = H.synth(1500, seed=66, reservoir=0, **H.CONFIGS["cfg3_320k_js_ms"])       # 1.5 MB: ~96 segments
    t = parallel_false_chain(s)
    info = compare_hop(ctx, t)
    assert info["n_frames"] == 1500 and info["rounds"] > 1, info
    info = compare_hop(ctx, t[7:])                                  # entering in the middle of a frame: the false header comes first, and wins -- on the host as well
    assert info["n_frames"] >= 1498


@pytest.mark.parametrize("mode", ["fast", "exact"])
@pytest.mark.parametrize("name", ["cfg3", "cfg4", "mono", "k48", "crc", "garbage", "nores"])
def test_decode_from_raw_bytes(name, mode):
    import pdmp3_b200
    s, _ = H.synth(260, seed=67, **VARIANTS[name])
    c = pdmp3_b200.Context(0, pdmp3_b200.MODE_FAST if mode == "fast" else pdmp3_b200.MODE_EXACT)
    a = c.decode(s, lookahead=1152)
    c.reset(); b, info = c.decode_raw(s, lookahead=1152)
    assert info["n_frames"] == a.shape[0] and np.array_equal(a, b)
    # streaming: batches of 61 frames, parser state and reservoir carried on the device side
    c.reset()
    st = pdmp3_b200._binding.P3ParseState(0, 0, 0, -1, -1)
    pos, parts = 0, []
    while True:
        p, info = c.decode_raw(s[pos:], lookahead=0, max_frames=61, state=st)
        if info["n_frames"] == 0: break
        parts.append(p); pos += info["consumed"]
        st.pcm_index = 0
    c.reset(); whole = c.decode(s, lookahead=0)
    c.close()
    assert np.array_equal(np.concatenate(parts), whole)


def test_mixing_host_parsed_and_raw_batches(ctx):
    """the reservoir carry moves between its host copy and its device copy as the two staging paths alternate"""
    import pdmp3_b200
    s, _ = H.synth(200, seed=68, **H.CONFIGS["cfg3_320k_js_ms"])
    ctx.reset(); whole = ctx.decode(s, lookahead=0)
    ctx.reset()
    st = pdmp3_b200._binding.P3ParseState(0, 0, 0, -1, -1)
    pos, parts, k = 0, [], 0
    while True:
        if k % 2 == 0:
            p, info = ctx.decode_raw(s[pos:], lookahead=0, max_frames=37, state=st)
            n, used = info["n_frames"], info["consumed"]
        else:
            pr = pdmp3_b200.parse_stream(s[pos:], lookahead=0, max_frames=37, state=st)
            n, used = pr.n_frames, pr.consumed
            for f in range(n): pr.c.frames[f].pcm_index = f
            p = ctx.decode_parsed(pr) if n else None
        if n == 0: break
        parts.append(p); pos += used; st.pcm_index = 0; k += 1
    assert np.array_equal(np.concatenate(parts), whole)


def test_one_million_frames_hop(ctx):
    """BASELINE configs[2] size: the 1 M-frame stream bench.py decodes (15 625-frame block x 64, 1.04 GB, 63 781 segments)"""
    import bench
    blk, _ = H.synth(bench.BLOCK, seed=1, **bench.CFG)
    s = np.tile(blk, 64)
    info = compare_hop(ctx, s)
    assert info["n_frames"] == 1000000 and info["rounds"] == 1
